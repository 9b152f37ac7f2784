"""The attenuation the loop produces against the analytic quality factor of its SLS set — the
pin of the anelastic path that does not go through the oracle's own arithmetic (tests/q_pin.py).

CPU: the oracle, coarse-grained memory variables; shear and bulk attenuation together and one at
a time; and the approach to the reference's `q_linear_solid` as dt -> 0.
GPU: the CUDA product library, coarse-grained and full memory variables."""
import os

import numpy as np
import pytest

from tests import q_pin

NRANKS = min(8, os.cpu_count() or 1)


def _oracle_run(qmu, qka, anel, cg=True, courant=0.6):
    from axisem_b200.capi import TimeLoop, connect_local, run_group
    from oracle import oracle
    probs, niter, dt = q_pin.problems(qmu, qka, anel, cg, NRANKS, courant)
    lib = oracle.load_fast()
    loops = [TimeLoop(lib, p) for p in probs]
    if NRANKS > 1:
        connect_local(lib, loops)
        run_group(lib, loops, niter)
    else:
        loops[0].run(niter)
    return q_pin.gather(probs, loops, niter), dt


@pytest.fixture(scope="module")
def elastic():
    return _oracle_run(40.0, 80.0, False)


@pytest.mark.parametrize("qmu,qka", [(40.0, 80.0), (40.0, 1.0e5), (1.0e5, 60.0)])
def test_oracle_attenuates_p_waves_with_the_q_of_its_sls_set(elastic, qmu, qka):
    """Q_P from the two-station spectral ratio within 5 % (energy-weighted over the source band)
    of the SLS set's, the dispersion within 5 %: with both mechanisms, with shear attenuation only
    (1/Q_P = L / Q_mu, L = 4/3 (vs/vp)^2) and with bulk attenuation only (1/Q_P = (1 - L) / Q_kappa)."""
    s_el, dt = elastic
    s_an, dt2 = _oracle_run(qmu, qka, True)
    assert dt == dt2
    r_im, r_re = q_pin.compare(s_el, s_an, dt, qmu, qka, discrete=True)
    print(f"Q_mu {qmu:g} Q_kappa {qka:g}: attenuation measured / SLS theory {r_im:.4f}, dispersion {r_re:.4f}")
    assert abs(r_im - 1.0) < 0.05, r_im
    assert abs(r_re - 1.0) < 0.05, r_re


def test_the_continuous_q_is_approached_linearly_in_dt(elastic):
    """Against q_linear_solid itself (the continuous SLS response) the loop attenuates too much,
    in proportion to dt: the reference forms the anelastic stress of step n+1 from the memory
    variables of step n (time_evol_wave.F90:395-455).  Halving dt halves the excess; extrapolated
    to dt -> 0 the measured Q is within 6 % of q_linear_solid's."""
    qmu, qka = 40.0, 80.0
    s_el, dt = elastic
    s_an, _ = _oracle_run(qmu, qka, True)
    r1, _ = q_pin.compare(s_el, s_an, dt, qmu, qka, discrete=False)
    s_el2, dth = _oracle_run(qmu, qka, False, courant=0.3)
    s_an2, _ = _oracle_run(qmu, qka, True, courant=0.3)
    r2, _ = q_pin.compare(s_el2, s_an2, dth, qmu, qka, discrete=False)
    print(f"excess attenuation over q_linear_solid: {r1 - 1:.3f} at dt = {dt:.3f} s, {r2 - 1:.3f} at dt = {dth:.3f} s; "
          f"extrapolated {2 * r2 - r1 - 1:.3f}")
    assert abs(dth - dt / 2) < 1e-12
    assert 0.15 < r1 - 1 < 0.40 and 0.08 < r2 - 1 < 0.25, (r1, r2)
    assert 0.45 < (r2 - 1) / (r1 - 1) < 0.75, (r1, r2)
    assert abs(2 * r2 - r1 - 1.0) < 0.06, 2 * r2 - r1


def test_complex_modulus_reproduces_q_linear_solid():
    """The continuous modulus used above gives exactly the reference's q_linear_solid(exact)."""
    from axisem_b200.host import AttenuationModel
    from axisem_b200.host.precomp import fast_correct
    att = AttenuationModel()
    w = 2 * np.pi * np.logspace(-3, 0, 40)
    for Q in (40.0, 312.0):
        ka, mu = q_pin.complex_moduli(att, Q, Q, w)
        q_mod = np.real(mu) / np.imag(mu)
        q_ref = q_pin.q_linear_solid(fast_correct(np.asarray(att.y_j) / Q), np.asarray(att.w_j), w)
        assert np.allclose(q_mod, q_ref, rtol=1e-12)
        # and the (synthetic default) set does what such a fit is for: Q within 12 % of the target over the band
        band = (w > 2 * np.pi * 0.003) & (w < 2 * np.pi * 0.5)
        assert np.all(np.abs(q_ref[band] / Q - 1) < 0.12), (q_ref[band] / Q).round(3)


@pytest.mark.gpu
@pytest.mark.parametrize("cg", [True, False])
def test_cuda_attenuates_p_waves_with_the_q_of_its_sls_set(cg):
    from axisem_b200 import solver
    qmu, qka = 40.0, 80.0
    out = {}
    for anel in (False, True):
        probs, niter, dt = q_pin.problems(qmu, qka, anel, cg, 1)
        loop = solver.time_loop(probs[0])
        loop.run(niter)
        out[anel] = q_pin.gather(probs, [loop], niter)
        assert loop.gpu_launches > 0
    r_im, r_re = q_pin.compare(out[False], out[True], dt, qmu, qka, discrete=True)
    print(f"{'cg4' if cg else 'full'} memory variables: attenuation measured / SLS theory {r_im:.4f}, dispersion {r_re:.4f}")
    assert abs(r_im - 1.0) < 0.05, r_im
    assert abs(r_re - 1.0) < 0.05, r_re
