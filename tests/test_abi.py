"""The C-ABI library loads on a CPU-only box, exports every symbol include/helmholtz_b200.h declares, its
host-side helpers agree with the oracle, and every compute entry point fails loudly without a GPU
(no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, rel_err


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "helmholtz_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hh_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/helmholtz_b200.h but not exported"
    # and the ctypes table binds exactly the header
    assert sorted(pkg._lib.SIGNATURES) == names
    assert lib.hh_version() == 100
    assert lib.hh_profile_num_tags() > 10 and lib.hh_profile_tag_name(0) == b"fine_apply"


def test_struct_layouts_match_header(pkg):
    L = pkg._lib
    assert C.sizeof(L.hh_mg_options) == 6 * 4 + 2 * 12 * 4 + 8 + 12 * 8
    assert C.sizeof(L.hh_solve_options) == 4 * 4 + 8


def test_host_helpers_match_oracle(pkg, ho):
    for n, neu, pad in (([19, 13], True, [4, 3]), ([19, 13], False, [4, 3]), ([11, 9, 10], True, [3, 2, 4]),
                        ([11, 9, 10], False, [3, 2, 4]), ([9, 7], False, [6, 5])):
        assert rel_err(pkg.getABL(n, neu, pad, 2.5), ho.getABL(n, neu, pad, 2.5)) < 1e-14
    mesh = pkg.getRegularMesh([0.0, 13.5, 0.0, 4.2], [256, 128])
    omesh = ho.getRegularMesh([0.0, 13.5, 0.0, 4.2], [256, 128])
    m = np.random.default_rng(0).uniform(0.1, 0.5, (257, 129))
    assert pkg.getMaximalFrequency(m, mesh) == pytest.approx(ho.getMaximalFrequency(m, omesh), rel=1e-15)
    assert pkg.loc2cs([257, 129], [128, 1]) == ho.loc2cs([257, 129], [128, 1]) == 128
    assert pkg.loc2cs([5, 4, 3], [2, 3, 2]) == ho.loc2cs([5, 4, 3], [2, 3, 2])
    q, src = pkg.getAcousticPointSource(mesh)
    qo, srco = ho.getAcousticPointSource(omesh)
    assert src == srco and np.array_equal(q.ravel(order="F"), qo)
    with pytest.raises(pkg._lib.HelmholtzB200Error):
        pkg.getABL([9, 7], True, [10, 2], 1.0)  # pad larger than the grid


def test_no_cpu_fallback(pkg):
    lib = pkg._lib.load()
    cnt = C.c_int(-1)
    assert lib.hh_device_count(C.byref(cnt)) == 0
    if cnt.value > 0:
        pytest.skip("a GPU is visible: the fallback check only makes sense on the CPU box")
    mesh = pkg.getRegularMesh([0, 1, 0, 1], [8, 8])
    with pytest.raises(pkg._lib.HelmholtzB200Error) as e:
        pkg.GetHelmholtzOperator(mesh, np.ones((9, 9)), 1.0, np.zeros((9, 9)), True, True)
    assert e.value.code == pkg._lib.HH_ERR_CUDA and "no CPU fallback" in str(e.value)
    # NULL-handle guards of every handle-taking entry point return an argument error, never crash
    assert lib.hh_setup(None, None) == pkg._lib.HH_ERR_ARG
    assert lib.hh_solve(None, None, None, 1, None, None, None) == pkg._lib.HH_ERR_ARG
    assert lib.hh_destroy(None) == 0
