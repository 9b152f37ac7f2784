"""Known-answer tests of the host-side inputs (not gpu).

Values are those of SURVEY.md Appendix A, derived from the reference formulas
(MESHER/splib.f90, MESHER/gllmeshgen.f90) for npol = 4; the remaining checks are the
reference's own self-checks: mass = volume (SOLVER/def_grid.f90:1188), integral of the
solid-fluid boundary term = 2 per boundary (SOLVER/def_precomp_terms.f90:2663, :2743).
"""
import numpy as np
import pytest

from axisem_b200.host import SpectralBasis, build_problem, prem_mesh_spec, SourceParams
from axisem_b200.host.mesh import build_rank
from axisem_b200.host.precomp import geometry


def test_gll_nodes_and_weights_npol4():
    b = SpectralBasis(4)
    np.testing.assert_allclose(b.eta, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=1e-14)
    np.testing.assert_allclose(b.wt, [0.1, 49 / 90, 32 / 45, 49 / 90, 0.1], atol=1e-14)


def test_glj01_nodes_and_weights_npol4():
    b = SpectralBasis(4)
    np.testing.assert_allclose(b.xi_k, [-1, -0.5077876295, 0.1323008208, 0.7088201422, 1], atol=2e-10)
    np.testing.assert_allclose(b.wt_axial_k, [0.0133333333, 0.2896566946, 0.7360043695,
                                              0.7943389360, 0.1666666667], atol=2e-10)
    assert abs(b.wt_axial_k.sum() - 2.0) < 1e-13          # = int (1 + xi) d xi


def test_derivative_matrices():
    b = SpectralBasis(4)
    for D, x in ((b.G2_dp, b.eta), (b.G1_dp, b.xi_k)):
        # D[j, i] = l_j'(x_i): exact for polynomials up to degree 4
        for p in range(5):
            np.testing.assert_allclose((x ** p) @ D, p * x ** max(p - 1, 0) if p else 0 * x, atol=1e-12)
    # stored in single precision, G2T/G1T are the transposes, G0 = G1(:,0)
    assert b.G2.dtype == np.float32 and b.G1.dtype == np.float32
    assert np.array_equal(b.G2T, b.G2.T) and np.array_equal(b.G1T, b.G1.T)
    assert np.array_equal(b.G0, b.G1[:, 0])


@pytest.mark.parametrize("ntheta", [4, 16])
def test_mass_matrix_sums_to_volume(ntheta):
    spec = prem_mesh_spec(ntheta=ntheta, nr_target=12)
    mesh = build_rank(spec, 0, 1)
    vol = 0.0
    for es in (mesh.solid, mesh.fluid):
        vol += 2 * np.pi * geometry(es, mesh.basis).massmat_k.sum()
    r0, r1 = spec.r_edges[0], spec.r_edges[-1]
    exact = 4.0 / 3.0 * np.pi * (r1 ** 3 - r0 ** 3)
    # theta is integrated by 5-point quadrature of sin: converges fast with ntheta
    assert abs(vol - exact) / exact < (2e-6 if ntheta == 4 else 1e-12)


def test_solid_fluid_boundary_term_integrates_to_two_per_boundary():
    spec = prem_mesh_spec(ntheta=16, nr_target=12)
    prob = build_problem(spec, SourceParams(), niter=4)
    m = prob.mesh
    assert m.nel_bdry == 2 * spec.ntheta                       # ICB and CMB
    es = m.solid
    sub_th = geometry(es, m.basis).th[m.bdry_solid_el - 1]     # (nb, 5)
    r = geometry(es, m.basis).r[m.bdry_solid_el - 1, m.bdry_jpol_solid]
    sign = np.where(m.bdry_above, 1.0, -1.0)
    B1, B2 = prob.bdry_matr[0].astype(np.float64), prob.bdry_matr[1].astype(np.float64)
    # B1 = +-r^2 dtheta w sin^2, B2 = +-r^2 dtheta w sin cos  ->  dtheta w sin
    integrand = (B1 * np.sin(sub_th) + B2 * np.cos(sub_th)) / (sign * r * r)[:, None]
    # axial point: the reference accumulates 1/r dtheta w_0 ds/dxi there (:2593), which is
    # what bdry_matr(0,iel,2)/r^2 holds (with cos = 1 at both poles, :2603)
    ax = es.axis[m.bdry_solid_el - 1]
    integrand[ax, 0] = (B2 / (sign * r * r)[:, None])[ax, 0]
    assert abs(integrand.sum() - 4.0) < 1e-5


def test_index_maps_are_consistent():
    spec = prem_mesh_spec(ntheta=8, nr_target=12)
    m = build_rank(spec, 0, 1)
    for nel, ig, nglob, gid in ((m.nel_solid, m.igloc_solid, m.nglob_solid, m.gid_solid),
                                (m.nel_fluid, m.igloc_fluid, m.nglob_fluid, m.gid_fluid)):
        assert ig.dtype == np.int32 and ig.min() == 1 and ig.max() <= nglob
        # local numbers and mesh-global ids induce the same partition of the points
        _, a = np.unique(ig, return_inverse=True)
        _, b = np.unique(gid, return_inverse=True)
        assert np.array_equal(a, b)
        # interior 3x3 points of an element are never shared (commun.F90:110-120)
        cnt = np.bincount(ig)[ig].reshape(nel, 5, 5)
        assert (cnt[:, 1:4, 1:4] == 1).all()
    # axis flags <-> ax_el lists (def_grid.f90:59-77)
    assert np.array_equal(np.nonzero(m.axis_solid)[0] + 1, np.sort(m.ax_el_solid))
    assert np.array_equal(np.nonzero(m.axis_fluid)[0] + 1, np.sort(m.ax_el_fluid))


@pytest.mark.parametrize("nranks", [2, 4])
def test_halo_lists_match_between_ranks(nranks):
    spec = prem_mesh_spec(ntheta=8, nr_target=12)
    meshes = [build_rank(spec, r, nranks) for r in range(nranks)]
    for dom in ("solid", "fluid"):
        for r, m in enumerate(meshes):
            h = getattr(m, "halo_" + dom)
            gid = getattr(m, "gid_" + dom)
            ig = getattr(m, "igloc_" + dom)
            first = {}
            for p, g in enumerate(ig):
                first.setdefault(int(g), p)
            assert h.nmsg <= 2 and set(h.list_peer.tolist()) <= {r - 1, r + 1}
            for k in range(h.nmsg):
                peer = meshes[int(h.list_peer[k])]
                hp = getattr(peer, "halo_" + dom)
                kk = list(hp.list_peer).index(r)
                assert hp.sizemsg[kk] == h.sizemsg[k]
                # the i-th entry of both lists is the same physical point
                mine = [gid[first[int(g)]] for g in h.glocal_index_msg[k, :h.sizemsg[k]]]
                pg, pig = getattr(peer, "gid_" + dom), getattr(peer, "igloc_" + dom)
                pfirst = {}
                for p, g in enumerate(pig):
                    pfirst.setdefault(int(g), p)
                theirs = [pg[pfirst[int(g)]] for g in hp.glocal_index_msg[kk, :hp.sizemsg[kk]]]
                assert mine == theirs
            # glob2el lists every local copy of every shared point
            pts = h.glob2el[:, 2].astype(np.int64) * 0
            if h.num_comm_gll:
                ipt = (h.glob2el[:, 2] - 1) * 25 + h.glob2el[:, 1] * 5 + h.glob2el[:, 0]
                shared = set(int(g) for k in range(h.nmsg) for g in h.glocal_index_msg[k, :h.sizemsg[k]])
                assert set(ig[ipt].tolist()) == shared
                assert np.isin(ig, list(shared)).sum() == h.num_comm_gll
