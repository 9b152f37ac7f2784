"""End-to-end pin against the reference's own golden seismograms:
TESTING/nightly/test_0{1,2,3}/ref_data/axisem.mseed — explosion (monopole), mtr (dipole), mtp
(quadrupole) in elastic prem_ani, the traces the reference's nightly regression compares every
build against — committed as tests/golden/nightly_ref_seismograms.npz.

See tests/nightly_compare.py for what is compared and why the agreement is at waveform level
(correlation, amplitude ratio) rather than sample level: the meshes differ, and source and
receivers sit on the nearest GLL points of either mesh.

CPU: the oracle (8 theta-slices in threads) on a coarse mesh, dipole source, all of E, N, Z.
GPU: the CUDA library on a finer mesh, all three source types."""
import os

import numpy as np
import pytest

from axisem_b200.host import SourceParams, build_problem, prem_mesh_spec
from tests.nightly_compare import compare, stations, to_enz

T_0 = 70.0          # Gaussian source period [s]: keeps the comparison inside the band both meshes resolve


def _setup(src, ntheta, nr, nranks, scheme="newmark2"):
    names, lat, lon = stations()
    colat = 90.0 - lat
    spec = prem_mesh_spec(ntheta=ntheta, nr_target=nr, anisotropic=True, r_min_km=800.0)
    sp = SourceParams(src_type2=src, depth=100e3, magnitude=1e20, t_0=T_0)
    dt = build_problem(spec, sp, niter=4, rec_colat_deg=colat, time_scheme=scheme).deltat
    shift = np.ceil(1.5 * T_0 / dt) * dt
    niter = int((1800.0 + shift) / dt) + 1
    seis_it = max(1, int(0.8 / dt))
    probs = [build_problem(spec, sp, niter=niter, rec_colat_deg=colat, seis_it=seis_it, rank=r, nranks=nranks,
                           time_scheme=scheme) for r in range(nranks)]
    return probs, niter, seis_it, dt, shift, np.deg2rad(colat), np.deg2rad(lon)


def _score(src, loops, probs, niter, seis_it, dt, shift, colat, lon, against="axisem"):
    ns = max(L.nseismo for L in loops)
    s = np.zeros((ns, colat.size, 3))
    for p, L in zip(probs, loops):
        if p.num_rec:
            s[:, p.rec_index, :] = L.seismograms()
    t = np.arange(ns) * seis_it * dt - shift                 # origin time = centre of the Gaussian
    r = compare(src, to_enz(src, s, colat, lon), t, T_0, against=against)
    big = r[:, :, 3] > 0.002 * r[:, :, 3].max()              # traces that carry signal (explosion: E is zero)
    return r[:, :, 0][big], r[:, :, 2][big], r


def test_oracle_reproduces_the_references_dipole_seismograms():
    from axisem_b200.capi import TimeLoop, connect_local, run_group
    from oracle import oracle
    nranks = min(8, os.cpu_count() or 1)
    probs, niter, *rest = _setup("mtr", 128, 40, nranks)
    lib = oracle.load_fast()
    loops = [TimeLoop(lib, p) for p in probs]
    if nranks > 1:
        connect_local(lib, loops)
        run_group(lib, loops, niter)
    else:
        loops[0].run(niter)
    cc, amp, _ = _score("mtr", loops, probs, niter, *rest)
    assert cc.size >= 40                                      # of 20 stations x (E, N, Z)
    assert cc.min() > 0.80 and np.median(cc) > 0.97, (cc.min(), np.median(cc))
    assert amp.min() > 0.80 and amp.max() < 1.20, (amp.min(), amp.max())


@pytest.mark.gpu
@pytest.mark.parametrize("src", ["explosion", "mtr", "mtp"])
def test_cuda_reproduces_the_references_seismograms(src):
    from axisem_b200 import solver
    probs, niter, *rest = _setup(src, 224, 60, 1)
    loop = solver.time_loop(probs[0])
    loop.run(niter)
    cc, amp, r = _score(src, [loop], probs, niter, *rest)
    print(f"{src}: {cc.size} traces, correlation min {cc.min():.4f} median {np.median(cc):.4f}, "
          f"amplitude ratio {amp.min():.3f} .. {amp.max():.3f}, launches {loop.gpu_launches}")
    assert loop.gpu_launches > 0
    assert cc.size >= 35
    assert cc.min() > 0.85 and np.median(cc) > 0.98, (cc.min(), np.median(cc))
    assert amp.min() > 0.70 and amp.max() < 1.30, (amp.min(), amp.max())
    if src == "explosion":
        # the independent YSPEC solution (full sphere, no attenuation, no gravity) that the
        # reference ships next to its own traces: test_01/ref_data/yspec.mseed
        cy, ay, _ = _score(src, [loop], probs, niter, *rest, against="yspec")
        print(f"  against yspec: {cy.size} traces, correlation min {cy.min():.4f} median {np.median(cy):.4f}, "
              f"amplitude ratio {ay.min():.3f} .. {ay.max():.3f}")
        assert cy.size >= 35 and cy.min() > 0.85 and np.median(cy) > 0.98, (cy.min(), np.median(cy))
        assert ay.min() > 0.70 and ay.max() < 1.30, (ay.min(), ay.max())


@pytest.mark.gpu
def test_cuda_symplectic_scheme_reproduces_the_references_dipole_seismograms():
    """The 4th-order symplectic loop (time step 1.5 x Newmark's, point-wise source time function
    of compute_stf_t) against the same golden traces; the oracle gives the same numbers."""
    from axisem_b200 import solver
    probs, niter, *rest = _setup("mtr", 128, 40, 1, scheme="symplec4")
    loop = solver.time_loop(probs[0])
    loop.run(niter)
    cc, amp, _ = _score("mtr", [loop], probs, niter, *rest)
    print(f"symplec4 mtr: {cc.size} traces, correlation min {cc.min():.4f} median {np.median(cc):.4f}, "
          f"amplitude ratio {amp.min():.3f} .. {amp.max():.3f}")
    assert cc.size >= 40
    assert cc.min() > 0.80 and np.median(cc) > 0.97, (cc.min(), np.median(cc))
    assert amp.min() > 0.80 and amp.max() < 1.20, (amp.min(), amp.max())
