"""The C-ABI shared libraries load and export every symbol include/axisem_b200.h declares;
without a GPU the product fails loudly instead of falling back (not gpu)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "axisem_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"AXB\((\w+)\)\s*\(", src)))


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    from axisem_b200.capi import SYMBOLS
    assert names == sorted(SYMBOLS)
    assert {"create", "run", "set_halo", "fetch_seismograms", "ipc_export"} <= set(names)


@pytest.mark.parametrize("lib,prefix", [("axisem_b200/libaxisem_b200.so", "axb_"),
                                        ("axisem_b200/libaxisem_b200_strict.so", "axb_"),
                                        ("oracle/libaxisem_oracle.so", "axo_")])
def test_library_exports_every_declared_symbol(lib, prefix):
    path = os.path.join(ROOT, lib)
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    syms = exported(path)
    missing = [prefix + n for n in declared_symbols() if prefix + n not in syms]
    assert not missing, missing
    C.CDLL(path)                                  # dlopen succeeds (CUDA runtime is linked statically)


def test_product_exports_nothing_but_the_abi():
    syms = exported(os.path.join(ROOT, "axisem_b200", "libaxisem_b200.so"))
    extra = [s for s in syms if not s.startswith("axb_")]
    assert not extra, extra


def test_product_does_not_reference_the_oracle():
    """the product path must not link, load or call anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "axisem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libaxisem_oracle" not in txt and "axo_" not in txt, f
    out = subprocess.run(["ldd", os.path.join(ROOT, "axisem_b200", "libaxisem_b200.so")],
                         capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the loud-failure path cannot be exercised")
    from axisem_b200 import solver
    from axisem_b200.capi import AxbError
    from tests.util import make_problem
    prob = make_problem("explosion", ntheta=4, nr=8, niter=4)
    with pytest.raises(AxbError, match="no CUDA device"):
        solver.time_loop(prob)


def test_bad_arguments_are_reported_like_the_reference_stops():
    """error behaviour of the boundary: non-zero return + message (the reference `stop`s)"""
    from oracle import oracle
    from axisem_b200.capi import AxbError, TimeLoop
    from tests.util import make_problem
    prob = make_problem("explosion", ntheta=4, nr=8, niter=4)
    loop = oracle.make_loop(prob)
    loop.run(4)
    with pytest.raises(AxbError, match="niter"):
        loop.run(1)                                # beyond niter
    with pytest.raises(AxbError):
        loop.seismograms(first=0, n=99)            # beyond the recorded samples


def test_integration_guide_binds_every_entry_point():
    """INTEGRATION.md's iso_c_binding module declares an interface for every symbol of the header, no more."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    syms = set(re.findall(r"AXB\((\w+)\)\(", open(os.path.join(root, "include", "axisem_b200.h")).read()))
    bound = set(re.findall(r"name='axb_(\w+)'", open(os.path.join(root, "INTEGRATION.md")).read()))
    assert syms == bound, (sorted(syms - bound), sorted(bound - syms))
