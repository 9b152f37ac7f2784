"""dump_type strain_only / fullfields (compute_strain, time_evol_wave.F90:1264-1410;
dump_velo_global, wavefields_io.f90:932-1015).

CPU: the oracle's dump against the Voigt strain of the attenuation module formed independently
in float64 (tests/test_oracle_physics.py: E12, E13, E23 of the dump are half the engineering
shears there), the packing of the ibeg:iend x jbeg:jend block, and the velocity fields.
GPU: the device kernels (axb_dump_fields.cuh) against the oracle — strict build bit-identical,
product build within 1e-5."""
import numpy as np
import pytest

from axisem_b200.host import SourceParams, build_problem
from tests.test_oracle_physics import _grad, _over_s, _voigt_strain
from tests.util import apply_state, rel_l2, seeded_state, small_spec

SRCS = ["explosion", "mtr", "mtp"]


def _problem(src, dump_type, block=(0, 4, 0, 4), niter=10, strain_it=5, anel=False):
    return build_problem(small_spec(), SourceParams(src_type2=src, t_0=40.0), niter=niter, dump=True,
                         strain_it=strain_it, dump_type=dump_type, dump_block=block, anel=anel)


def _expected_solid(prob, src, u):
    E = _voigt_strain(prob, u, src)                       # (nel, 6, j, i): E1 E2 E3 E4 E5 E6
    out = {"strain_dsus": E[:, 0], "strain_dpup": E[:, 1], "strain_dsuz": E[:, 4] / 2,
           "straintrace": E[:, 0] + E[:, 1] + E[:, 2]}
    if src != "explosion":
        out["strain_dsup"] = E[:, 5] / 2
        out["strain_dzup"] = E[:, 3] / 2
    if src == "mtp":
        # The two strain routines of the reference disagree for the quadrupole: compute_strain takes
        # E12 = -(us + up/2)/s - ds(up)/2 (time_evol_wave.F90:1341-1343), compute_strain_att_el
        # E6/2 = (up/2 - us)/s - ds(up)/2 (attenuation.f90:596-600) — the sign of up/(2s) differs.
        # The dump restates compute_strain as written.
        m, b = prob.mesh, prob.mesh.basis
        pw = {k: v.astype(np.float64) for k, v in prob.pw_solid.items()}
        ax = m.axis_solid.astype(bool)
        u1, u2, _ = u.astype(np.float64)
        out["strain_dsup"] = -_over_s(u1 + u2 / 2, pw, ax, b) - _grad(u2, pw, ax, b)[0] / 2
    return out


def _names(src, full):
    n = ["strain_dsus", "strain_dsuz", "strain_dpup"] + ([] if src == "explosion" else ["strain_dsup", "strain_dzup"]) \
        + ["straintrace"]
    if full:
        n += ["velo_s"] + ([] if src == "explosion" else ["velo_p"]) + ["velo_z"]
    return n


@pytest.mark.parametrize("src", SRCS)
def test_oracle_strain_dump_is_the_strain_tensor(src):
    from oracle import oracle
    prob = _problem(src, "fullfields", (0, 4, 0, 4))
    O = oracle.make_loop(prob)
    st = seeded_state(O, scale=1e-3, fields=("disp", "velo", "chi", "dchi"))
    apply_state(O, st)
    O.run(1)                                   # the dump at iter 0 sees the seeded state
    snap = O.snapshots()[:, 0, :]
    names = _names(src, True)
    assert snap.shape[0] == len(names)
    m = prob.mesh
    ns, nf = m.nel_solid, m.nel_fluid
    assert snap.shape[1] == 25 * (ns + nf)
    sol = {n: snap[k, :25 * ns].reshape(ns, 5, 5) for k, n in enumerate(names)}
    flu = {n: snap[k, 25 * ns:].reshape(nf, 5, 5) for k, n in enumerate(names)}
    want = _expected_solid(prob, src, st["disp"])
    for n, w in want.items():
        assert rel_l2(sol[n], w) < 2e-5, (n, rel_l2(sol[n], w))
    v = st["velo"].astype(np.float64)
    if src == "mtr":
        assert np.array_equal(sol["velo_s"], st["velo"][0] + st["velo"][1])
        assert np.array_equal(sol["velo_p"], st["velo"][0] - st["velo"][1])
    else:
        assert np.array_equal(sol["velo_s"], st["velo"][0])
    assert np.array_equal(sol["velo_z"], st["velo"][2])
    # fluid: u = grad(chi) / rho, then the same strain operators; phi components from u / s
    pw = {k: a.astype(np.float64) for k, a in prob.pw_fluid.items()}
    ax = m.axis_fluid.astype(bool)
    b = m.basis
    ir = prob.inv_rho_fluid.astype(np.float64)
    gs, gz = _grad(st["chi"].astype(np.float64), pw, ax, b)
    us, uz = gs * ir, gz * ir
    a1, a2 = _grad(us, pw, ax, b)
    b1, b2 = _grad(uz, pw, ax, b)
    fs, fz = _over_s(us, pw, ax, b), _over_s(uz, pw, ax, b)
    wantf = {"strain_dsus": a1, "strain_dsuz": (a2 + b1) / 2, "strain_dpup": fs, "straintrace": fs + a1 + b2}
    if src == "mtr":
        wantf.update(strain_dsup=-fs / 2, strain_dzup=fz / 2)
    elif src == "mtp":
        wantf.update(strain_dsup=-fs, strain_dzup=-fz)
    for n, w in wantf.items():
        assert rel_l2(flu[n], w) < 5e-5, (n, rel_l2(flu[n], w))
    ws, wz = _grad(st["dchi"].astype(np.float64), pw, ax, b)
    assert rel_l2(flu["velo_s"], ws * ir) < 2e-5 and rel_l2(flu["velo_z"], wz * ir) < 2e-5
    if src != "explosion":
        assert not flu["velo_p"].any()


def test_oracle_fullfields_block_and_strain_only_mapping():
    """fullfields packs floc(ibeg:iend, jbeg:jend, :) in Fortran order; strain_only uses the kwf
    mapping (de-duplicated points): both are selections of the same 25-point fields."""
    from oracle import oracle
    full = {}
    for key, dump_type, block in (("all", "fullfields", (0, 4, 0, 4)), ("blk", "fullfields", (1, 3, 1, 3)),
                                  ("kwf", "strain_only", (0, 4, 0, 4))):
        prob = _problem("mtr", dump_type, block)
        O = oracle.make_loop(prob)
        apply_state(O, seeded_state(O, scale=1e-3, fields=("disp", "velo", "chi", "dchi")))
        O.run(6)                               # dumps at iter 0 and 5
        assert O.nstrain == 2
        full[key] = (O.snapshots(), prob)
    a, prob = full["all"]
    m = prob.mesh
    ns, nf = m.nel_solid, m.nel_fluid
    blk = full["blk"][0]
    assert blk.shape[2] == 9 * (ns + nf)
    for v in range(9):
        for s in range(2):
            fs_ = a[v, s, :25 * ns].reshape(ns, 5, 5)[:, 1:4, 1:4]
            ff_ = a[v, s, 25 * ns:].reshape(nf, 5, 5)[:, 1:4, 1:4]
            assert np.array_equal(blk[v, s, :9 * ns].reshape(ns, 3, 3), fs_)
            assert np.array_equal(blk[v, s, 9 * ns:].reshape(nf, 3, 3), ff_)
    kwf = full["kwf"][0]
    q = prob.kwf
    mask = q["kwf_mask"].reshape(-1).astype(bool)
    idx = q["mapping_ijel_ikwf"].reshape(-1)[mask] - 1
    for v in range(6):
        for s in range(2):
            assert np.array_equal(kwf[v, s, idx], a[v, s, mask])


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("dump_type,block", [("strain_only", (0, 4, 0, 4)), ("fullfields", (0, 4, 0, 4)),
                                             ("fullfields", (1, 3, 0, 4))])
@pytest.mark.parametrize("src", SRCS)
def test_cuda_field_dumps_match_the_oracle(src, dump_type, block, strict):
    from axisem_b200 import solver
    from oracle import oracle
    n = 40
    prob = _problem(src, dump_type, block, niter=n, strain_it=8, anel=True)
    G, O = solver.time_loop(prob, strict=strict), oracle.make_loop(prob)
    st = seeded_state(G, scale=1e-9)
    for L in (G, O):
        apply_state(L, st)
        L.run(n // 2)
        L.run(n - n // 2)
    assert G.nstrain == O.nstrain == n // 8 + 1
    g, o = G.snapshots(), O.snapshots()
    assert g.shape == o.shape and np.abs(o).max() > 0
    for v in range(o.shape[0]):
        if strict:
            assert np.array_equal(g[v], o[v]), (v, rel_l2(g[v], o[v]))
        elif np.abs(o[v]).max() > 0:
            assert rel_l2(g[v], o[v]) <= 1e-5, (v, rel_l2(g[v], o[v]))
    assert np.array_equal(G.seismograms(), O.seismograms()) if strict else rel_l2(G.seismograms(), O.seismograms()) <= 1e-5
