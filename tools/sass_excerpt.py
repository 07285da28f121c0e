"""Trimmed SASS evidence for profiles/: per hot kernel the instruction count, the opcode histogram and the memory /
reduction / tensor-free proof lines (LDG widths, LDGSTS, IDP.4A, REDUX, RED, ATOM), plus a window of the hot loop.
usage: python tools/sass_excerpt.py telomeric_identifier_b200/libtidk_cuda.so profiles/r02_sass_excerpts.md"""
import collections, re, subprocess, sys

lib, out = sys.argv[1], sys.argv[2]
KERNELS = [("window_counts_flat_kernel<StaticMatcher<6,TTAGGG>,3>", r"window_counts_flat_kernelINS_13StaticMatcherILi6ELm2703EEELi3"),
           ("window_counts_flat_kernel<FlatRuntime<12>,3>", r"window_counts_flat_kernelINS_11FlatRuntimeILi12EEELi3"),
           ("window_counts_kernel<StaticMatcher<6,TTAGGG>,4,AsciiSource>", r"window_counts_kernelINS_13StaticMatcherILi6ELm2703EEELi4ENS_11AsciiSource"),
           ("explore_vertical_kernel<5,12>", r"explore_vertical_kernelILi5ELi12"),
           ("planes_build_kernel", r"planes_build_kernel")]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
with open(out, "w") as f:
    f.write("# SASS excerpts (sm_100a), `cuobjdump -sass` of the shipped `libtidk_cuda.so`\n\n"
            "No tensor-core instruction (`UTC*MMA`, `HMMA`, `IMMA`) appears in any kernel of the library: the path is integer streaming "
            "work (loads, LOP3 / SHF / PRMT / POPC, IDP.4A, REDUX, RED).  TMA bulk copies (`UBLKCP`) with mbarrier completion (`SYNCS`) "
            "stage the plane words of the vertical explore scan; the window counters read with plain 128- / 256-bit loads (nothing is "
            "reused across lanes there, DESIGN 3.1).\n\n")
    allops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", txt, re.M))
    f.write("Whole library, opcode families of interest: " + ", ".join(f"`{k}` x{allops.get(k, 0)}" for k in
            ("LDG", "LDGSTS", "IDP", "LOP3", "SHF", "PRMT", "POPC", "REDUX", "RED", "REDG", "ATOMG", "SHFL", "HMMA", "IMMA", "UTCHMMA", "UBLKCP", "SYNCS")) + "\n\n")
    for name, pat in KERNELS:
        body = next((x for x in funcs if re.match(r"_ZN?4?\w*" , x) and re.search(pat, x.split("\n", 1)[0])), None)
        if body is None:
            f.write(f"## {name}\n\nnot found in this build\n\n")
            continue
        lines = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l)]
        ins = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() for l in lines]
        ops = collections.Counter(re.sub(r"^/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?", "", l).split()[0].split(".")[0] for l in ins)
        f.write(f"## {name}\n\n{len(ins)} instructions ({len(ins) * 16 // 1024} KB). Opcodes: " +
                ", ".join(f"{k} {v}" for k, v in ops.most_common(14)) + "\n\n```\n")
        keep = [l for l in ins if re.search(r"\b(LDG|LDGSTS|STG|IDP|REDUX|RED|REDG|ATOMG|ATOM|SHFL|CCTL|PREFETCH|UBLKCP|SYNCS)\b|LDG\.|RED\.|ATOMG\.|IDP\.|REDUX\.|UBLKCP|SYNCS\.", l)]
        for l in keep[:40]:
            f.write(l[:130] + "\n")
        f.write("```\n\nHot-loop window (first 48 instructions after the kernel's first LOP3-dense stretch):\n\n```\n")
        dense = next((i for i in range(len(ins) - 24) if sum("LOP3" in x or "SHF" in x for x in ins[i:i + 24]) >= 18), 0)
        for l in ins[dense:dense + 48]:
            f.write(l[:130] + "\n")
        f.write("```\n\n")
print("wrote", out)
