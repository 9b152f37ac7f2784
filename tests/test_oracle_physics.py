"""Pins of the CPU oracle that do not need the Fortran (not gpu).

The reference cannot be built here and its tests hold no array-level vectors for the time
loop, so the oracle is pinned by:

  * the strain-energy identity  1/2 u.K u = int W(strain(u)) s ds dz, where K is the
    restated glob_stiffness_{mono,di,quad}_4 / glob_fluid_stiffness_4 with the restated
    def_precomp_terms planes, and the strain is formed independently in float64 from the
    reference's own strain definition (attenuation.f90:542-606: Voigt strain per source
    order) and the isotropic energy density.  A wrong sign, plane or index in either the
    pre-computation or the stiffness routines breaks it;
  * symmetry of K (also with TI anisotropy and on axial elements);
  * energy conservation of the Newmark loop (the reference's SAVE_ENERGY diagnostic,
    time_evol_wave.F90:1424);
  * equality of seismograms between 1 rank and 2/4 theta-slices (commun.F90/commpi.F90).
"""
import numpy as np
import pytest

from axisem_b200.capi import connect_local, run_group
from axisem_b200.host import SourceParams, build_problem, homogeneous_layers, prem_mesh_spec
from axisem_b200.host.mesh import MeshSpec
from axisem_b200.host.precomp import geometry, material
from oracle import oracle
from tests.util import make_problem, rel_l2

SRCS = ["explosion", "mtr", "mtp"]


def _axis_bc(u, src, ax):
    """fields that satisfy the axial boundary conditions the masks enforce"""
    if src == "explosion":
        u[0][ax, :, 0] = 0
        u[1] = 0
    elif src == "mtr":
        u[1][ax, :, 0] = 0
        u[2][ax, :, 0] = 0
    else:
        u[:, ax, :, 0] = 0
    return u


def _grad(f, pw, ax, b):
    G2, G1 = b.G2.astype(np.float64), b.G1.astype(np.float64)       # G[j,i] = l_j'(x_i)
    dxi = np.where(ax[:, None, None], np.einsum("ejk,ki->eji", f, G1), np.einsum("ejk,ki->eji", f, G2))
    deta = np.einsum("eki,kj->eji", f, G2)
    ds = pw["DzDeta_over_J"] * dxi + pw["DzDxi_over_J"] * deta
    dz = pw["DsDeta_over_J"] * dxi + pw["DsDxi_over_J"] * deta
    return ds, dz


def _over_s(f, pw, ax, b):
    r = f * pw["inv_s"]
    r[ax, :, 0] = _grad(f, pw, ax, b)[0][ax, :, 0]        # l'Hospital on the axis
    return r


def _solid_strain_energy(prob, u, src):
    m, b = prob.mesh, prob.mesh.basis
    g = geometry(m.solid, b)
    _, lam, mu, *_ = material(m.spec, m.solid, g)
    pw = {k: v.astype(np.float64) for k, v in prob.pw_solid.items()}
    ax = m.axis_solid.astype(bool)
    u1, u2, u3 = u.astype(np.float64)
    gr = lambda f: _grad(f, pw, ax, b)
    fs = lambda f: _over_s(f, pw, ax, b)
    b1 = gr(u1 + u2) if src == "mtr" else gr(u1)
    b2 = gr(u3)
    E = [b1[0], None, b2[1], 0 * u1, b1[1] + b2[0], 0 * u1]
    if src == "explosion":
        E[1] = fs(u1)
    elif src == "mtr":
        E[1] = 2 * fs(u2)
        c = gr(u1 - u2)
        E[3] = -fs(u3) - c[1]
        E[5] = -E[1] - c[0]
    else:
        E[1] = fs(u1 - 2 * u2)
        c = gr(u2)
        E[3] = -2 * fs(u3) - c[1]
        E[5] = fs(u2 - 2 * u1) - c[0]
    tr = E[0] + E[1] + E[2]
    W = 0.5 * lam * tr ** 2 + mu * (E[0] ** 2 + E[1] ** 2 + E[2] ** 2) \
        + 0.5 * mu * (E[3] ** 2 + E[4] ** 2 + E[5] ** 2)
    return (g.massmat_k * W).sum()


@pytest.mark.parametrize("src", SRCS)
def test_solid_stiffness_equals_strain_energy(src):
    prob = make_problem(src, anisotropic=False, ntheta=8, nr=10)
    O = oracle.make_loop(prob)
    rng = np.random.default_rng(1)
    ax = prob.mesh.axis_solid.astype(bool)
    for _ in range(3):
        u = _axis_bc(rng.standard_normal(O._field_shape("disp")).astype(np.float32), src, ax)
        O.set("disp", u)
        O.apply_op("solid_stiffness")
        a = 0.5 * (u * O.get("acc1").astype(np.float64)).sum()
        e = _solid_strain_energy(prob, u, src)
        assert abs(a / e - 1.0) < 2e-7, (src, a, e)


@pytest.mark.parametrize("src,m", [("explosion", 0), ("mtr", 1), ("mtp", 2)])
def test_fluid_stiffness_equals_potential_energy(src, m):
    prob = make_problem(src, ntheta=8, nr=10)
    O = oracle.make_loop(prob)
    mesh, b = prob.mesh, prob.mesh.basis
    g = geometry(mesh.fluid, b)
    rho = material(mesh.spec, mesh.fluid, g)[0]
    pw = {k: v.astype(np.float64) for k, v in prob.pw_fluid.items()}
    ax = mesh.axis_fluid.astype(bool)
    rng = np.random.default_rng(2)
    for _ in range(3):
        c = rng.standard_normal(O._field_shape("chi")).astype(np.float32)
        if m:
            c[ax, :, 0] = 0
        O.set("chi", c)
        O.apply_op("fluid_stiffness")
        a = 0.5 * (c * O.get("ddchi1").astype(np.float64)).sum()
        f = c.astype(np.float64)
        ds, dz = _grad(f, pw, ax, b)
        W = 0.5 / rho * (ds ** 2 + dz ** 2 + (m * _over_s(f, pw, ax, b)) ** 2)
        e = (g.massmat_k * W).sum()
        assert abs(a / e - 1.0) < 2e-7, (src, a, e)


@pytest.mark.parametrize("src", SRCS)
def test_solid_stiffness_is_symmetric_with_anisotropy(src):
    prob = make_problem(src, anisotropic=True, ntheta=8, nr=10)
    O = oracle.make_loop(prob)
    rng = np.random.default_rng(3)
    ax = prob.mesh.axis_solid.astype(bool)

    def K(u):
        O.set("disp", u)
        O.apply_op("solid_stiffness")
        return O.get("acc1").astype(np.float64)

    u = _axis_bc(rng.standard_normal(O._field_shape("disp")).astype(np.float32), src, ax)
    v = _axis_bc(rng.standard_normal(O._field_shape("disp")).astype(np.float32), src, ax)
    a, b = (v * K(u)).sum(), (u * K(v)).sum()
    assert abs(a - b) / abs(a) < 5e-6


def _unique_weights(ig):
    """1 for the first local copy of every global point, 0 for the others"""
    w = np.zeros(ig.size)
    w[np.unique(ig, return_index=True)[1]] = 1.0
    return w


def test_newmark_conserves_energy_in_a_solid_sphere():
    """Elastic, source-free, monopole: E = 1/2 v.M v + 1/2 u.K u stays constant
    (time_evol_wave.F90:1424-1526 is the reference's version of this diagnostic)."""
    spec = MeshSpec(ntheta=8, layers=homogeneous_layers(), nrad=[8])
    prob = build_problem(spec, SourceParams(src_type2="explosion", magnitude=0.0), niter=400,
                         rec_colat_deg=[30.0])
    O = oracle.make_loop(prob)
    m = prob.mesh
    assert m.nel_fluid == 0
    # a smooth initial displacement (continuous across elements): u_s ~ s f(r), u_z ~ z f(r)
    _, _, r, s, z, *_ = m.coords("solid")
    R = np.broadcast_to(r[:, :, None], s.shape)
    f = np.exp(-((R - 3.5e6) / 8e5) ** 2) * 1e-7
    u0 = np.zeros(O._field_shape("disp"), np.float32)
    u0[0], u0[2] = s * f, z * f
    O.set("disp", u0)
    w = _unique_weights(m.igloc_solid).reshape(m.nel_solid, 5, 5)
    mass = 1.0 / prob.inv_mass_rho.astype(np.float64)

    def energy():
        u, v = O.get("disp").astype(np.float64), O.get("velo").astype(np.float64)
        P = oracle.make_loop(prob)
        P.set("disp", u.astype(np.float32))
        P.apply_op("solid_stiffness")
        pot = 0.5 * (u * P.get("acc1").astype(np.float64)).sum()
        kin = 0.5 * (w * mass * (v[0] ** 2 + v[2] ** 2)).sum()
        return pot + kin, pot, kin

    e0 = energy()[0]
    es = []
    for _ in range(8):
        O.run(50)
        es.append(energy())
    tot = np.array([e[0] for e in es])
    assert max(e[2] for e in es) > 0.05 * e0              # energy really moves into motion
    assert np.abs(tot / e0 - 1.0).max() < 5e-3, tot / e0  # Newmark: bounded O(dt^2) wobble
    assert abs(tot[-1] / tot[0] - 1.0) < 2e-3             # no drift


@pytest.mark.parametrize("src,anel", [("explosion", False), ("mtr", True), ("mtp", False)])
@pytest.mark.parametrize("nranks", [2, 4])
def test_theta_slices_reproduce_single_rank(src, anel, nranks):
    n = 60
    kw = dict(anel=anel, ntheta=8, nr=12, niter=n, t_0=8.0)
    one = oracle.make_loop(make_problem(src, **kw))
    one.run(n)
    ref = one.seismograms()
    probs = [make_problem(src, rank=r, nranks=nranks, **kw) for r in range(nranks)]
    loops = [oracle.make_loop(p) for p in probs]
    lib = oracle.load()
    connect_local(lib, loops)
    run_group(lib, loops, n)
    got = np.zeros_like(ref)
    seen = np.zeros(ref.shape[1], bool)
    for p, L in zip(probs, loops):
        got[:, p.rec_index] = L.seismograms()
        seen[p.rec_index] = True
    assert seen.all()
    assert np.abs(ref).max() > 0
    # only the order of the halo sums differs (commpi.F90:469-477): a few ulp per step
    assert rel_l2(got, ref) < 1e-5


# ---------------------------------------------------------------------------------------
# Anelastic K term: virtual work.  glob_anel_stiffness_* subtracts B^T R, so for any test
# field v and memory variables R (Voigt, engineering shear)  v . l(R) = int R : E(v) s ds dz
# with E(v) the strain of compute_strain_att_el_* formed independently in float64.
def _voigt_strain(prob, u, src):
    m, b = prob.mesh, prob.mesh.basis
    pw = {k: v.astype(np.float64) for k, v in prob.pw_solid.items()}
    ax = m.axis_solid.astype(bool)
    u1, u2, u3 = u.astype(np.float64)
    gr = lambda f: _grad(f, pw, ax, b)
    fs = lambda f: _over_s(f, pw, ax, b)
    b1 = gr(u1 + u2) if src == "mtr" else gr(u1)
    b2 = gr(u3)
    E = [b1[0], None, b2[1], 0 * u1, b1[1] + b2[0], 0 * u1]
    if src == "explosion":
        E[1] = fs(u1)
    elif src == "mtr":
        E[1] = 2 * fs(u2)
        c = gr(u1 - u2)
        E[3] = -fs(u3) - c[1]
        E[5] = -E[1] - c[0]
    else:
        E[1] = fs(u1 - 2 * u2)
        c = gr(u2)
        E[3] = -2 * fs(u3) - c[1]
        E[5] = fs(u2 - 2 * u1) - c[0]
    return np.stack(E, axis=1)                       # (nel, 6, jpol, ipol)


def _anel_work(prob, src, cg, v, rng, u):
    O = oracle.make_loop(prob)
    R = np.zeros(O._field_shape("memvar"), np.float32)
    R[:, 1, v] = rng.standard_normal(R[:, 1, v].shape)      # one SLS, one Voigt component
    R[:, 3, v] = rng.standard_normal(R[:, 3, v].shape)      # and a second one: they add up
    O.set("memvar", R)
    O.set("acc1", np.zeros_like(u))
    O.apply_op("anel_stiffness")
    lhs = -(u * O.get("acc1").astype(np.float64)).sum(axis=(0, 2, 3))
    g = geometry(prob.mesh.solid, prob.mesh.basis)
    E = _voigt_strain(prob, u, src)[:, v]
    Rs = R.astype(np.float64).sum(axis=1)[:, v]
    if cg:
        w = np.stack([g.massmat_k[:, 1, 1], g.massmat_k[:, 3, 1], g.massmat_k[:, 1, 3], g.massmat_k[:, 3, 3]], axis=1)
        Ek = np.stack([E[:, 1, 1], E[:, 3, 1], E[:, 1, 3], E[:, 3, 3]], axis=1)
        rhs = (w * Rs * Ek).sum(axis=1)
    else:
        rhs = (g.massmat_k * Rs * E).sum(axis=(1, 2))
    return lhs, rhs


@pytest.mark.parametrize("cg", [True, False])
@pytest.mark.parametrize("src", SRCS)
def test_anelastic_stiffness_is_the_adjoint_of_the_strain(src, cg):
    prob = make_problem(src, anel=True, coarse_grained=cg, ntheta=8, nr=10)
    ax = prob.mesh.axis_solid.astype(bool)
    rng = np.random.default_rng(5)
    u = _axis_bc(rng.standard_normal((3,) + prob.pw_solid["inv_s"].shape).astype(np.float32), src, ax)
    for v in ([0, 1, 2, 4] if src == "explosion" else range(6)):
        lhs, rhs = _anel_work(prob, src, cg, v, rng, u)
        scale = np.abs(rhs[~ax]).max()
        assert np.abs(lhs - rhs)[~ax].max() <= 5e-6 * scale, (src, cg, v)


@pytest.mark.parametrize("src", SRCS)
def test_full_memvar_axial_terms_are_adjoint_once_the_reference_s_quirk_is_removed(src):
    """In axial elements the reference evaluates s of V_* at eta(ipol) instead of xi_k(ipol)
    (def_precomp_terms.f90:1485-1496; reproduced by host/precomp.py).  With s put back at the
    GLJ point the Y0/V0_* axial branch of glob_anel_stiffness_*_4 is the exact adjoint of the
    L'Hopital strain for every component the reference's axial branch treats completely
    (dipole r4/r5 and quadrupole r4/r5 are not: stiffness_di.f90:708-711,
    stiffness_quad.f90:645-656)."""
    prob = make_problem(src, anel=True, coarse_grained=False, ntheta=8, nr=10)
    m, b = prob.mesh, prob.mesh.basis
    ax = m.axis_solid.astype(bool)
    g = geometry(m.solid, b)
    th_q = 0.5 * ((1.0 - b.eta[None, :]) * m.solid.th_a[:, None] + (1.0 + b.eta[None, :]) * m.solid.th_b[:, None])
    s_q = g.r[:, :, None] * np.sin(th_q)[:, None, :]
    ratio = np.ones_like(g.s)
    ok = ax[:, None, None] & (s_q > 0)
    ratio[ok] = g.s[ok] / s_q[ok]
    for n in ("V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi"):
        prob.solid[n] = (prob.solid[n].astype(np.float64) * ratio).astype(np.float32)
    rng = np.random.default_rng(6)
    u = _axis_bc(rng.standard_normal((3,) + prob.pw_solid["inv_s"].shape).astype(np.float32), src, ax)
    for v in ([0, 1, 2, 4] if src == "explosion" else [0, 1, 2, 5]):
        lhs, rhs = _anel_work(prob, src, False, v, rng, u)
        scale = np.abs(rhs[ax]).max()
        assert np.abs(lhs - rhs)[ax].max() <= 1e-4 * scale, (src, v)


@pytest.mark.parametrize("src", ["explosion", "mtp"])
def test_reference_energy_diagnostic_is_conserved(src):
    """dump_energy (time_evol_wave.F90:1424-1526) restated in the oracle: after the source has
    died away, (epot + ekin) of solid + fluid stays constant in the coupled PREM-type model —
    which pins the restated stiffness, mass and S/F coupling against each other."""
    from axisem_b200.host import prem_mesh_spec
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    prob = build_problem(spec, SourceParams(src_type2=src, t_0=40.0), niter=1200, energy=True)
    over = int((1.5 * 40.0 + 40.0) / prob.deltat)          # gauss_0: shift 1.5 t_0, width t_0 / 3.5
    n = over + 130
    assert n <= 1200
    O = oracle.make_loop(prob)
    O.run(n)
    e = O.energy().astype(np.float64)
    assert e.shape == (n + 1, 4) and np.all(e[0] == 0.0)
    tot = 0.5 * e.sum(axis=1) * 2 * np.pi
    assert tot[over] > 0
    assert np.abs(tot[over:] / tot[over] - 1.0).max() < 2e-4


def test_reference_energy_diagnostic_dipole_quirk():
    """For dipole sources the reference weights the z kinetic term with two * (two * m)
    (def_precomp_terms.f90:749-750 and time_evol_wave.F90:1484), while the loop advances u_z with
    the mass m (inv_mass_rho carries 1/2 and the z corrector a factor 2, :479): the diagnostic as
    written is not a conserved quantity; with the z term weighted by m it is.  The oracle and the
    device restate the reference's formula as it is."""
    from axisem_b200.host import prem_mesh_spec
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    prob = build_problem(spec, SourceParams(src_type2="mtr", t_0=40.0), niter=1200, energy=True)
    O = oracle.make_loop(prob)
    um = prob.unassem_mass_rho_solid.astype(np.float64)
    over = int((1.5 * 40.0 + 40.0) / prob.deltat)
    n = over + 132
    assert n <= 1200
    O.run(over)
    ref, fixed = [], []
    for _ in range(4):
        O.run((n - over) // 4)
        e = O.energy(O.iter, 1)[0].astype(np.float64)
        v = O.get("velo").astype(np.float64)
        kz = 2.0 * (v[2] ** 2 * um).sum()
        ref.append(e.sum())
        fixed.append(e.sum() - kz + kz / 4.0)
    ref, fixed = np.array(ref), np.array(fixed)
    assert np.abs(fixed / fixed[0] - 1.0).max() < 2e-4
    assert np.abs(ref / ref[0] - 1.0).max() > 1e-2
