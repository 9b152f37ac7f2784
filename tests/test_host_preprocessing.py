"""The first native pieces of the pre-processing that sits in front of the seam (SURVEY.md
section 8f item 1, second half): spectral basis (MESHER/splib.f90, gllmeshgen.f90:61-94) and
background models (SOLVER/background_models.F90:417-674) in C++, via axisem_b200_hosttool.

The background model is pinned against the reference's own tabulation of prem_ani
(tests/golden/prem_ani_model_bm.npz, from TESTING/TEST04_anelastic_anisotropic/model.bm); the
spectral basis against the known values of SURVEY.md Appendix A and the Python builder."""
import os
import subprocess

import numpy as np
import pytest

from axisem_b200.capi import fortran_matrix
from axisem_b200.host.spectral import SpectralBasis

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_hosttool")


def _exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["bash", os.path.join(ROOT, "axisem_b200", "hostcxx", "build.sh")])
    return EXE


def _spectral(npol):
    out = subprocess.run([_exe(), "spectral", str(npol)], capture_output=True, text=True, check=True).stdout
    return {l.split()[0]: np.array([float(v) for v in l.split()[1:]]) for l in out.splitlines()}


def test_spectral_basis_known_answers_npol4():
    s = _spectral(4)
    r37 = np.sqrt(3.0 / 7.0)
    assert np.allclose(s["eta"], [-1, -r37, 0, r37, 1], atol=1e-15)
    assert np.allclose(s["wt"], [0.1, 49 / 90, 32 / 45, 49 / 90, 0.1], atol=1e-15)
    assert abs(s["wt"].sum() - 2.0) < 1e-15
    assert abs(s["wt_axial_k"].sum() - 2.0) < 1e-14          # integral of (1 + xi) over [-1, 1]
    assert np.allclose((s["wt_axial_k"] * s["xi_k"]).sum(), 2.0 / 3.0, atol=1e-14)
    n1 = 5
    for name, x in (("G2", s["eta"]), ("G1", s["xi_k"])):
        G = s[name].reshape(n1, n1).T                          # Fortran order [j + n1*i] -> G[j, i] = l_j'(x_i)
        assert np.allclose(G.sum(axis=0), 0.0, atol=2e-6)      # derivative of a constant
        assert np.allclose((G * x[:, None]).sum(axis=0), 1.0, atol=2e-6)   # derivative of x
        assert np.allclose(s[name + "T"].reshape(n1, n1), s[name].reshape(n1, n1).T)
    assert np.array_equal(s["G0"], s["G1"].reshape(n1, n1)[0])        # G0(j) = G1(j, 0)


@pytest.mark.parametrize("npol", [4, 5, 6])
def test_spectral_basis_matches_the_python_builder(npol):
    s = _spectral(npol)
    if npol == 4:
        b = SpectralBasis(4)
    else:
        from axisem_b200.host.spectral import gll_points_weights, glj_points_weights, lagrange_deriv_matrix
        b = SpectralBasis.__new__(SpectralBasis)
        b.eta, b.wt = gll_points_weights(npol)
        b.xi_k, b.wt_axial_k = glj_points_weights(npol)
        b.G2 = lagrange_deriv_matrix(b.eta).astype(np.float32)
        b.G1 = lagrange_deriv_matrix(b.xi_k).astype(np.float32)
        b.G2T, b.G1T = b.G2.T, b.G1.T
        b.G0 = b.G1[:, 0]
    for k in ("eta", "wt", "xi_k", "wt_axial_k"):
        assert np.allclose(s[k], getattr(b, k), rtol=0, atol=5e-15), k
    for k in ("G1", "G1T", "G2", "G2T"):
        # real(4) values: the two implementations may differ in the last bit of the real(8) source
        assert np.allclose(s[k], fortran_matrix(getattr(b, k)), rtol=2e-7, atol=1e-7), k
    assert np.allclose(s["G0"], b.G0, rtol=2e-7)


def test_native_prem_ani_matches_the_references_tabulated_model():
    z = np.load(os.path.join(ROOT, "tests", "golden", "prem_ani_model_bm.npz"))
    t = z["table"]
    lines = []
    for k, row in enumerate(t):
        upper = k == 0 or t[k - 1, 0] != row[0]               # a discontinuity radius appears twice, upper side first
        lines.append(f"{row[0]:.3f} {'u' if upper else 'l'}")
    out = subprocess.run([_exe(), "model", "prem_ani"], input="\n".join(lines) + "\n", capture_output=True,
                         text=True, check=True).stdout
    got = np.array([[float(v) for v in l.split()] for l in out.splitlines()])
    assert got.shape == (160, 9)
    tol = np.array([0.006, 0.006, 0.006, 0.006, 0.006, 6e-6, 0.5, 0.005])
    assert np.all(np.abs(got[:, :8] - t[:, 1:]) <= tol), np.abs(got[:, :8] - t[:, 1:]).max(axis=0)
    idom = got[:, 8].astype(int)
    assert idom[0] == 1 and idom[-1] == 12 and np.all(np.diff(idom) >= 0)      # surface inwards, 12 domains
    fluid = t[:, 3] == 0.0
    assert set(idom[fluid]) == {11}


def test_native_prem_iso_matches_the_python_layers():
    from axisem_b200.host.model import R_EARTH, prem_layers
    layers = prem_layers(anisotropic=False, r_min_km=0.0)
    radii, side = [], []
    for L in layers:
        for f in (0.0, 0.37, 1.0):
            radii.append(L.r_bot + f * (L.r_top - L.r_bot))
            side.append("u" if f < 1.0 else "l")               # bottom of a layer = upper side of the discontinuity
    inp = "\n".join(f"{r:.6f} {s}" for r, s in zip(radii, side)) + "\n"
    out = subprocess.run([_exe(), "model", "prem_iso"], input=inp, capture_output=True, text=True, check=True).stdout
    got = np.array([[float(v) for v in l.split()] for l in out.splitlines()])
    k = 0
    for L in layers:
        for f in (0.0, 0.37, 1.0):
            x = radii[k] / R_EARTH
            want = [L.rho(x) * 1e3, L.vpv(x) * 1e3, L.vsv(x) * 1e3, L.vph(x) * 1e3, L.vsh(x) * 1e3, float(L.eta(x)),
                    L.qkappa, L.qmu]
            assert np.allclose(got[k, :8], want, rtol=1e-13, atol=1e-9), (L.name, f, got[k], want)
            k += 1


def test_hosttool_errors():
    r = subprocess.run([_exe(), "model", "nosuchmodel"], input="1000.0 u\n", capture_output=True, text=True)
    assert r.returncode == 1 and "unknown background model" in r.stderr
    r = subprocess.run([_exe(), "spectral", "1"], capture_output=True, text=True)
    assert r.returncode == 1
