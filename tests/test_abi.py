"""CPU: the C-ABI library loads without a GPU and exports every symbol declared in include/*.h."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(edgpu_[a-z0-9_]+|ed_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(edb):
    L = edb.lib()
    decl = declared_symbols("edgpu.h") + declared_symbols("ed_b200.h")
    assert len(decl) > 50
    missing = [s for s in decl if not hasattr(L, s)]
    assert missing == []
    # the python binding table is in sync with the headers
    assert sorted(edb.EDGPU_SYMBOLS + edb.ED_SYMBOLS) == sorted(decl)


def test_no_cpu_fallback_without_gpu(edb):
    """Without a CUDA device edgpu_init must fail loudly (never silently compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(edb.EdgpuError, match="no CPU fallback"):
        edb.Context(1, 4)


def test_product_does_not_reference_oracle():
    """The product sources must never include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "dmft-ed_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cpp", ".h", ".py", ".sh")):
                txt = open(os.path.join(dp, f)).read()
                assert "ed_oracle" not in txt and "oracle/" not in txt, f


def test_input_defaults_match_reference(edb):
    inp = edb.default_input()
    # ED_INPUT_VARS.f90:121-196
    assert (inp.Norb, inp.Nbath, inp.Nspin) == (1, 6, 1)
    assert list(inp.uloc)[:3] == [2.0, 0.0, 0.0]
    assert (inp.beta, inp.xmu, inp.hfmode) == (1000.0, 0.0, 1)
    assert (inp.Lmats, inp.Lreal, inp.wini, inp.wfin, inp.eps) == (5000, 5000, -5.0, 5.0, 0.01)
    assert (inp.lanc_niter, inp.lanc_ngfiter, inp.lanc_dim_threshold) == (512, 200, 256)
    assert inp.lanc_tolerance == 1e-12 and inp.gs_threshold == 1e-9
    assert edb.lib().ed_get_bath_dimension(ctypes.byref(inp)) == 12
