"""No-GPU checks of the boundary: the library loads, exports every symbol the header declares,
and refuses to run without a device (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tidk_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tidk_cuda_[a-z_]+)\s*\(", text)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    from telomeric_identifier_b200.api import EXPORTS, load_library
    L = load_library()
    syms = declared_symbols()
    assert len(syms) >= 13
    assert sorted(EXPORTS) == syms
    for s in syms:
        assert hasattr(L, s), s
    assert L.tidk_cuda_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    import telomeric_identifier_b200 as T
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    with pytest.raises(T.TidkCudaError) as e:
        T.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "telomeric_identifier_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower(), os.path.join(dp, f)
