"""theta x r decomposition (the mesher's NTHETA_SLICES x NRADIAL_SLICES, pdb.f90 / commpi.F90:
a rank then has up to 8 neighbours, the corner points are shared by 4 ranks): the host builder's
blocks, run by the oracle over `connect_local`, against the undivided run."""
import numpy as np
import pytest

from axisem_b200.capi import connect_local, run_group
from axisem_b200.host import SourceParams, build_problem, prem_mesh_spec
from axisem_b200.host.mesh import radial_blocks
from oracle import oracle
from tests.util import rel_l2

COLAT = np.linspace(5.0, 175.0, 9)


def _problems(spec, nranks, nranks_r, n, src="mtr", **kw):
    return [build_problem(spec, SourceParams(src_type2=src, t_0=3.0), anel=True, niter=n, rank=r, nranks=nranks,
                          nranks_r=nranks_r, rec_colat_deg=COLAT, **kw) for r in range(nranks)]


def test_radial_blocks_cover_the_radius_and_keep_the_sf_boundaries_inside():
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    for nr_r in (1, 2, 3, 4):
        blocks = radial_blocks(spec, nr_r)
        assert blocks[0][0] == 0 and blocks[-1][1] == spec.nr
        for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
            assert b == c and a < b
            # a cut never coincides with a solid/fluid boundary (the S/F coupling stays rank-local)
            assert spec.fluid_ir[b - 1] == spec.fluid_ir[b]


@pytest.mark.parametrize("nranks,nranks_r", [(2, 2), (4, 2), (6, 3), (8, 2)])
def test_blocks_equal_the_undivided_run(nranks, nranks_r):
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    n = 150
    one = build_problem(spec, SourceParams(src_type2="mtr", t_0=3.0), anel=True, niter=n, rec_colat_deg=COLAT)
    O1 = oracle.make_loop(one)
    O1.run(n)
    s1 = O1.seismograms()
    probs = _problems(spec, nranks, nranks_r, n)
    assert sum(p.mesh.nel_solid for p in probs) == one.mesh.nel_solid
    assert sum(p.mesh.nel_fluid for p in probs) == one.mesh.nel_fluid
    assert sum(p.nelsrc > 0 for p in probs) == 1 and sum(p.num_rec for p in probs) == COLAT.size
    if nranks // nranks_r > 2:
        assert max(p.mesh.halo_solid.nmsg for p in probs) == 5       # 2 columns, 1 row, 2 corners
    lib = oracle.load()
    loops = [oracle.make_loop(p) for p in probs]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    s = np.zeros_like(s1)
    for p, L in zip(probs, loops):
        if p.num_rec:
            s[:, p.rec_index, :] = L.seismograms()
    assert np.abs(s1).max() > 0
    assert rel_l2(s, s1) <= 1e-5
    # the wavefield itself, element by element
    u1 = O1.get("disp")
    for p, L in zip(probs, loops):
        m = p.mesh
        sel = np.nonzero((one.mesh.solid.it >= m.it0) & (one.mesh.solid.it < m.it1)
                         & (one.mesh.solid.ir >= m.ir0) & (one.mesh.solid.ir < m.ir1))[0]
        assert sel.size == m.nel_solid
        assert rel_l2(L.get("disp"), u1[:, sel] if u1.shape[1] == one.mesh.nel_solid else u1[sel]) <= 1e-5
