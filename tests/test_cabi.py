"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/viorb_gpu.h declares; without a GPU it fails loudly instead of falling back."""
import os
import re

import pytest

from util import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "viorb_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(viorb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from viorb_b200 import build
    path = build.build_cuda()
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libviorb_b200.so does not export %s" % s


def test_binding_covers_header():
    from viorb_b200 import api
    api.lib()
    assert set(declared_symbols()) <= set(api.EXPORTED)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from viorb_b200 import api
    with pytest.raises(api.ViorbError) as e:
        api.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may touch oracle/"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "viorb_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"orb_oracle|oracle_py|from oracle|import oracle|orc_[a-z]", txt) and f != "build.py":
                    bad.append(f)
    assert not bad, bad
