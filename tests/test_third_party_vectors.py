"""rust-bio 2.0.3 semantics the reference leans on without pinning them in its own tests (SURVEY.md 8c;
VERDICT round 1, "pin what the reference leaves unpinned"): KMP / BOM all-occurrence search incl.
self-overlapping hits, and bio::io::fasta::Reader's line handling.  The vectors
(tests/golden/third_party_vectors.json) are written from the crate's published doc-tests / unit tests and
its documented behaviour -- not from a run of the reference, which cannot be built here -- so the label
stays "parity unpinned"; what these tests establish is that the oracle, the independent Python model, the
host-side FASTA readers and (under -m gpu) the CUDA counters all agree with the same written-down rules."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pymodel as M
from telomeric_identifier_b200 import fasta

HERE = os.path.dirname(os.path.abspath(__file__))
V = json.load(open(os.path.join(HERE, "golden", "third_party_vectors.json")))


def bom_case(c):
    pat = c["unit"] * c["pattern_units"] + c["pattern_tail"]
    text = (c.get("text_prefix", "") + c["unit"] * c["text_units"] + c.get("text_mid", "")
            + c["unit"] * c.get("text_units2", 0) + c["text_tail"])
    return pat, text


def all_find_cases():
    for c in V["find_all"]:
        yield c["pattern"], c["text"], c["occ"]
    for c in V["bom_65"]:
        pat, text = bom_case(c)
        assert len(pat) >= 65  # utils.rs:25: the BOM branch
        yield pat, text, c["occ"]


@pytest.mark.parametrize("pat,text,occ", list(all_find_cases()))
def test_find_all_oracle_and_model(pat, text, occ):
    assert O.find_motifs(pat.encode(), text.encode()) == occ
    assert M.find_motifs(pat, text) == occ
    # naive definition, for the vectors themselves
    assert [i for i in range(len(text) - len(pat) + 1) if text[i:i + len(pat)] == pat] == occ


@pytest.mark.parametrize("pat,text,occ", list(all_find_cases()))
def test_window_counts_oracle_counts_every_occurrence(pat, text, occ):
    """search.rs:111-128: the window count is the length of the (overlapping) index list."""
    seq = np.frombuffer(text.encode(), dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    f, r = O.window_counts(seq, off, max(1, seq.size), [pat.encode()], True)
    assert f.tolist() == [len(occ)]


@pytest.mark.parametrize("c", V["fasta"], ids=[c["name"] for c in V["fasta"]])
def test_fasta_python_reader(c):
    ids, seq, off = fasta.parse_fasta(c["text"].encode())
    assert ids == c["ids"]
    assert [seq[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(ids))] == c["seqs"]


def test_fasta_python_reader_errors():
    for c in V["fasta_errors"]:
        with pytest.raises(ValueError):
            fasta.parse_fasta(c["text"].encode())


def fnv64(b: bytes) -> int:
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) % (1 << 64)
    return h


def test_fasta_cpp_reader(tmp_path):
    """The C++ driver's reader (csrc/host/fasta.hpp) on the same vectors, plain and gzipped, 1 and 4 threads."""
    import gzip
    import __graft_entry__ as g
    g.build()
    tidk = os.path.join(os.path.dirname(HERE), "telomeric_identifier_b200", "bin", "tidk")
    for c in V["fasta"]:
        for gz in (False, True):
            p = tmp_path / (c["name"] + (".fa.gz" if gz else ".fa"))
            data = c["text"].encode()
            p.write_bytes(gzip.compress(data) if gz else data)
            for threads in (1, 4):
                out = subprocess.run([tidk, "fasta-check", str(p), "--threads", str(threads)], capture_output=True, text=True)
                assert out.returncode == 0, (c["name"], out.stderr)
                rows = out.stdout.splitlines()
                total = "".join(c["seqs"]).encode()
                assert rows[0] == f"{len(c['ids'])}\t{len(total)}\t{fnv64(total):016x}", c["name"]
                assert rows[1:] == [f"{i}\t{len(s)}" for i, s in zip(c["ids"], c["seqs"])], c["name"]
    for c in V["fasta_errors"]:
        p = tmp_path / (c["name"] + ".fa")
        p.write_bytes(c["text"].encode())
        assert subprocess.run([tidk, "fasta-check", str(p)], capture_output=True, text=True).returncode == 1


def test_reference_binary_hook():
    """BASELINE.md section 4 reserves baseline/_ref/tidk for a real reference binary.  There is none in this
    image (no cargo / rustc); if one ever appears, the find_all vectors are diffed against it automatically."""
    exe = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "tidk")
    if not os.path.exists(exe):
        pytest.skip("no reference binary (the Rust reference cannot be built in this image)")
    import tempfile
    for pat, text, occ in all_find_cases():
        if not set(pat) <= set("ACGT"):
            continue
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "x.fa")
            open(fa, "w").write(">r\n" + text + "\n")
            subprocess.run([exe, "search", "-s", pat, "-w", str(len(text)), "-o", "x", "-d", d, fa], check=True)
            rows = open(os.path.join(d, "x_telomeric_repeat_windows.tsv")).read().splitlines()[1:]
            assert int(rows[0].split("\t")[2]) == len(occ)


@pytest.mark.gpu
@pytest.mark.parametrize("pat,text,occ", list(all_find_cases()))
def test_find_all_through_the_c_abi(pat, text, occ):
    import telomeric_identifier_b200 as T
    seq = np.frombuffer(text.encode(), dtype=np.uint8)
    off = np.array([0, seq.size], dtype=np.uint64)
    with T.Context(0) as ctx:
        f, r = ctx.window_counts(seq, off, max(1, seq.size), [pat.upper().encode()])
        of, orr = O.window_counts(seq, off, max(1, seq.size), [pat.encode()], True)
        assert f.tolist() == [len(occ)] == of.tolist()
        assert r.tolist() == orr.tolist()
