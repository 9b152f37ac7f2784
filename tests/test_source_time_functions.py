"""Every SOURCE_FUNCTION of the reference (inparam_advanced: errorf, gauss_0, gauss_1, gauss_2,
quheavi, dirac_0) on the host side and in the symplectic loop.

The formulas under test restate SOLVER/source.f90:144-233, 587-917 and parameters.F90:975-1072;
the checks are the analytic properties each function is defined by (unit integral of the Dirac
approximations, errorf' = gauss_0, the Heaviside as the running integral of the Dirac), the
agreement of the three restatements (numpy host, C++ host, C oracle) and the oracle's symplectic
loop against the Newmark loop on the same smooth source."""
import math
import os
import subprocess

import numpy as np
import pytest

from axisem_b200.capi import NSTAGES
from axisem_b200.host.source import (DIRAC_APPROX, SourceParams, compute_stf, compute_stf_t,
                                     discrete_dirac_setup, erf_nr, stf_shift)
from .util import make_problem

DT = 0.25


def test_erf_is_the_numerical_recipes_one():
    x = np.linspace(-5.0, 5.0, 2001)
    exact = np.array([math.erf(v) for v in x])
    err = np.abs(erf_nr(x) - exact).max()
    assert err < 1.3e-7          # the published accuracy of erfcc is 1.2e-7 everywhere
    assert err > 1e-9            # ... and it is that approximation, not the libm function


@pytest.mark.parametrize("stf", ["gauss_0", "gauss_1", "gauss_2", "errorf"])
def test_smooth_source_time_functions(stf):
    p = SourceParams(stf_type=stf, t_0=20.0, magnitude=3.0e19)
    n = 1200
    s = compute_stf(p, n, DT).astype(np.float64)
    t = np.arange(1, n + 1) * DT
    shift = stf_shift(p, DT)
    assert abs(shift / DT - round(shift / DT)) < 1e-9 and shift >= 1.5 * p.t_0
    if stf == "gauss_0":         # unit-area Gaussian times the magnitude, centred on the shift
        assert abs(s.sum() * DT / p.magnitude - 1.0) < 1e-6
        assert abs(t[s.argmax()] - shift) <= DT
    elif stf == "gauss_1":       # normalised to a peak of +- magnitude, odd about the shift
        assert abs(np.abs(s).max() / p.magnitude - 1.0) < 1e-3
        assert abs(s.sum()) * DT < 1e-6 * p.magnitude * p.t_0
    elif stf == "gauss_2":       # normalised by its side lobes 2 a^2 exp(-3/2): centre = -exp(3/2)/2
        k = int(round(shift / DT)) - 1
        assert abs(s[k] / p.magnitude + 0.5 * math.exp(1.5)) < 1e-5
    else:                        # moment function: 0 -> magnitude, derivative = gauss_0
        assert abs(s[0]) < 1e-6 * p.magnitude and abs(s[-1] / p.magnitude - 1.0) < 1e-6
        g = compute_stf(SourceParams(stf_type="gauss_0", t_0=20.0, magnitude=3.0e19), n, DT).astype(np.float64)
        d = np.gradient(s, DT)
        assert np.abs(d - g)[5:-5].max() < 2e-3 * g.max()


@pytest.mark.parametrize("choice", DIRAC_APPROX)
def test_discrete_dirac_and_quasi_heaviside(choice):
    half, _, shift = discrete_dirac_setup(period=40.0, deltat=DT, seis_it=4)
    p = SourceParams(stf_type="dirac_0", t_0=half, shift_seconds=shift, discrete_choice=choice, magnitude=2.0e20)
    n = 4000
    s = compute_stf(p, n, DT)
    assert s.dtype == np.float32
    assert abs(s.astype(np.float64).sum() * DT / p.magnitude - 1.0) < 1e-6     # delta_src normalises the sum
    if choice == "1dirac":
        assert np.count_nonzero(s) == 1 and np.nonzero(s)[0][0] + 1 == int(shift / DT)
    else:
        assert abs((s.argmax() + 1) * DT - shift) <= DT
    assert np.array_equal(s, compute_stf(SourceParams(**{**p.__dict__, "stf_type": "dirac_1"}), n, DT))
    h = compute_stf(SourceParams(**{**p.__dict__, "stf_type": "quheavi"}), n, DT).astype(np.float64)
    assert np.allclose(h, np.cumsum(s.astype(np.float64) * DT), rtol=1e-6, atol=0)
    assert abs(h[-1] / p.magnitude - 1.0) < 1e-6
    if choice in ("gaussi", "triang", "1dirac", "cauchy", "caulor"):
        assert (np.diff(h) >= -1e-7 * p.magnitude).all()                            # positive kernels: monotone


def test_discrete_dirac_setup_follows_parameters_f90():
    # seismograms at every step, no wavefield dumps: one-sample spike, half width period / 8
    half, choice, shift = discrete_dirac_setup(50.0, 0.5, seis_it=1)
    assert (half, choice) == (6.25, "1dirac") and shift == 0.5 * math.ceil(4 * 6.25 / 0.5) + 0.5
    # down-sampled seismograms: Gaussian of half width period / max(15, int(period / (2 seis_dt)))
    half, choice, shift = discrete_dirac_setup(50.0, 0.5, seis_it=4)
    assert choice == "gaussi" and half == float(np.float32(50.0 / 15))
    assert shift > 4 * half and abs(shift / 2.0 - round(shift / 2.0)) < 1e-6
    half, choice, shift = discrete_dirac_setup(200.0, 0.5, seis_it=4)
    assert half == float(np.float32(200.0 / 50))
    # ... for dirac_0 only (the reference compares with 'queavi' there): quheavi keeps period / 8
    half, choice, _ = discrete_dirac_setup(200.0, 0.5, seis_it=4, stf_type="quheavi")
    assert (half, choice) == (25.0, "gaussi")
    # wavefield dumps: Gaussian of half width period / 8, shift on the dump grid too
    half, choice, shift = discrete_dirac_setup(50.0, 0.5, seis_it=2, strain_it=8, dump_wavefields=True)
    assert (half, choice) == (6.25, "gaussi")
    for step in (0.5, 1.0, 4.0):
        assert abs(shift / step - round(shift / step)) < 1e-6
    assert shift >= 4 * half


def _ulp32(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


SYMP_KW = {
    "gauss_0": {}, "gauss_1": {}, "gauss_2": {}, "errorf": {},
    "dirac_0": {"shift_seconds": 2.0}, "quheavi": {"shift_seconds": 2.0},
}


def _coefd(scheme):
    """coefd(1:nstages) of symplectic_coefficients (time_evol_wave.F90:760-830), in units of deltat;
    the literals of symplec4 and ML_SO6m7 are default-real constants."""
    f = lambda v: float(np.float32(v))
    if scheme == "symplec4":
        zeta, kappa = f(0.1786178958448091), f(-0.06626458266981849)
        return np.array([zeta, kappa, 1.0 - 2.0 * (zeta + kappa), kappa])
    if scheme == "ML_SO4m5":
        rho, theta = (14.0 - math.sqrt(19.0)) / 108.0, (20.0 - 7.0 * math.sqrt(19.0)) / 108.0
        return np.array([rho, theta, 0.5 - rho - theta, 0.5 - rho - theta, theta])
    assert scheme == "ML_SO6m7"
    d = [f(-1.01308797891717472981), f(1.18742957373254270702), f(-0.01833585209646059034), f(0.34399425728109261313)]
    return np.array(d + d[::-1][:3])


@pytest.mark.parametrize("scheme", ["symplec4", "ML_SO4m5", "ML_SO6m7"])
@pytest.mark.parametrize("stf", sorted(SYMP_KW))
def test_oracle_symplectic_stf_table(stf, scheme):
    """compute_stf_t of the oracle at the sub-stage times of every step against the numpy
    restatement: time accumulated step by step, sub-stage offsets coeff(i) = sum coefd(1:i), the hat
    function switched by the first sub-stage, the index-based quasi-Heaviside."""
    from oracle import oracle
    n, seis_it = 60, 3
    prob = make_problem("mtr", niter=n, scheme=scheme, t_0=4.0, seis_it=seis_it,
                        source_kw={"stf_type": stf, "magnitude": 1.0e20, **SYMP_KW[stf]})
    O = oracle.make_loop(prob)
    tab = O.stf_symp(0, n)
    ns = NSTAGES[scheme]
    assert tab.shape == (n, ns)
    d = _coefd(scheme) * prob.deltat
    coeff = np.array([d[:k + 1].sum() for k in range(ns)])
    src = prob.source
    shift = stf_shift(src, prob.deltat)
    t = 0.0
    ref = np.zeros((n, ns))
    for it in range(n):
        t += prob.deltat
        ref[it] = compute_stf_t(src, t - prob.deltat + coeff, prob.deltat, seis_it)
    scale = np.abs(ref).max()
    assert scale > 0
    if stf == "quheavi":
        assert np.array_equal(tab, ref.astype(np.float32))
    else:          # libm exp against numpy's; summation order of coeff
        assert np.abs(tab - ref).max() <= 1e-6 * scale
    if stf == "dirac_0" and scheme == "symplec4":
        # a hat of width 2 deltat and height ~ magnitude / deltat around the shift, non-zero in the two
        # steps whose first sub-stage lies within deltat of the shift (the other sub-stages may lie
        # before the first one — kappa < 0 — so the peak can exceed 1 / deltat slightly; with the
        # large negative first coefficient of ML_SO6m7 it does so grossly: the reference's own
        # "do not work with symplectic schemes" remark in delta_src)
        assert 0.5 * src.magnitude / prob.deltat < tab.max() < 1.2 * src.magnitude / prob.deltat
        steps = np.nonzero(np.abs(tab).max(axis=1) > 0)[0] + 1
        k = int(round(shift / prob.deltat))
        assert set(steps) <= {k, k + 1, k + 2} and len(steps) == 2
    if stf == "quheavi":
        assert (tab[:, :seis_it - 1] == 0).all() and (tab[:, seis_it - 1:] == np.float32(src.magnitude)).all()
    O.close()


def test_oracle_rejects_nothing_silently():
    """dirac_1 has no case in compute_stf_t: the host refuses it for a symplectic scheme."""
    with pytest.raises(ValueError):
        compute_stf_t(SourceParams(stf_type="dirac_1"), np.array([0.0, 0.1]), 0.1)


@pytest.mark.parametrize("stf", ["errorf", "gauss_1"])
def test_symplectic_loop_with_a_smooth_source_follows_newmark(stf):
    """The two loops integrate the same equation: with the source time function sampled by
    compute_stf (Newmark) and by compute_stf_t (symplec4) the seismograms agree to the accuracy of
    the second-order scheme."""
    from oracle import oracle
    from .util import rel_l2
    from axisem_b200.host import build_problem
    from .util import small_spec
    n = 400
    src = SourceParams(src_type2="mtr", t_0=30.0, stf_type=stf, magnitude=1.0e20)
    a = build_problem(small_spec(), src, niter=n, time_scheme="newmark2")
    b = build_problem(small_spec(), src, niter=n, time_scheme="symplec4", deltat=a.deltat)
    assert np.abs(a.stf).max() > 0 and a.deltat == b.deltat and not b.stf.any()
    A, B = oracle.make_loop(a), oracle.make_loop(b)
    A.run(n)
    B.run(n)
    sa, sb = A.seismograms(), B.seismograms()
    assert np.abs(sa).max() > 0
    assert rel_l2(sb, sa) < 1e-3          # measured 1e-4
    A.close()
    B.close()
