"""XDMF snapshots (dump_xdmf: meshes_io.F90:110-464, wavefields_io.f90:119-738).

CPU: the plot-point maps of the host against their definition; the oracle's snapshot fields
(u, straintrace, curlinplane) against the same quantities formed independently in float64; the
files against the structure the reference's format statements prescribe.
GPU: the device kernels against the oracle — strict build bit-identical, product build 1e-5."""
import re
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from axisem_b200.host import SourceParams, build_problem
from axisem_b200.host.xdmf import write_xdmf, xdmf_maps
from tests.test_oracle_physics import _grad, _over_s
from tests.util import apply_state, rel_l2, seeded_state, small_spec

SRCS = ["explosion", "mtr", "mtp"]


def _problem(src, niter=12, snap_it=5, opts=None, **kw):
    return build_problem(small_spec(), SourceParams(src_type2=src, t_0=40.0), niter=niter, snap_it=snap_it,
                         xdmf_opts=opts, **kw)


def test_plot_point_maps():
    prob = _problem("mtr")
    m, x = prob.mesh, prob.xdmf
    nf, ns = m.nel_fluid, m.nel_solid
    mask, mp = x["plotting_mask"], x["mapping_ijel_iplot"]
    assert mask.shape == (nf + ns, 3, 3) and x["nelem_plot"] == 4 * (nf + ns)
    # every plot point has exactly one owner, numbered in visiting order, fluid first
    owners = mp[mask == 1]
    assert np.array_equal(np.sort(owners), np.arange(1, x["npoint_plot"] + 1))
    assert mp[:nf].max() < mp[nf:][mask[nf:] == 1].min()
    # points with the same global number share the plot point; fluid and solid never do
    ig = np.concatenate([m.igloc_fluid.reshape(nf, 5, 5), m.igloc_solid.reshape(ns, 5, 5) + m.nglob_fluid])
    sub = ig[:, ::2, ::2]
    for g in np.unique(sub)[:200]:
        assert np.unique(mp[sub == g]).size == 1
    assert np.unique(sub).size == x["npoint_plot"]
    # coordinates of the plot points and corner order of the cells (counter-clockwise in (i, j))
    _, th, r, s, z, _, _ = m.coords("solid")
    el, j, i = 3, 1, 2
    k = mp[nf + el, j, i] - 1
    assert np.allclose(x["points"][k], [s[el, 2 * j, 2 * i], z[el, 2 * j, 2 * i]], rtol=1e-6)
    c = x["grid"][4 * (nf + el)]                    # cell (i=0, j=0) of that element
    assert list(c) == [mp[nf + el, 0, 0] - 1, mp[nf + el, 0, 1] - 1, mp[nf + el, 1, 1] - 1, mp[nf + el, 1, 0] - 1]


def test_plot_region_limits_the_elements():
    full = _problem("explosion").xdmf
    part = _problem("explosion", opts=dict(rmin=5.0e6, thetamax=np.pi / 2)).xdmf
    assert 0 < part["nelem_plot"] < full["nelem_plot"] and part["npoint_plot"] < full["npoint_plot"]
    r = np.hypot(part["points"][:, 0], part["points"][:, 1])
    assert r.min() > 4.5e6                          # whole elements are kept: a little below rmin
    assert part["points"][:, 1].min() > -1.0e6


@pytest.mark.parametrize("src", SRCS)
def test_oracle_snapshot_fields(src):
    from oracle import oracle
    prob = _problem(src)
    O = oracle.make_loop(prob)
    st = seeded_state(O, scale=1e-3, fields=("disp", "velo", "chi", "dchi"))
    apply_state(O, st)
    O.run(1)                                        # snapshot 1 at iter 0 sees the seeded state
    snap = O.xdmf_snapshots()
    assert snap.shape == (5, 1, prob.xdmf["npoint_plot"])
    m, b, x = prob.mesh, prob.mesh.basis, prob.xdmf
    nf, ns = m.nel_fluid, m.nel_solid
    mask, mp = x["plotting_mask"].astype(bool), x["mapping_ijel_iplot"] - 1
    u = st["disp"].astype(np.float64)
    pw = {k: v.astype(np.float64) for k, v in prob.pw_solid.items()}
    ax = m.axis_solid.astype(bool)
    if src == "mtr":
        us, up, uz = u[0] + u[1], u[0] - u[1], u[2]
        phi = 2 * _over_s(u[1], pw, ax, b)
    elif src == "explosion":
        us, up, uz = u[0], u[1], u[2]
        phi = _over_s(u[0], pw, ax, b)
    else:
        us, up, uz = u[0], u[1], u[2]
        phi = _over_s(u[0] - 2 * u[1], pw, ax, b)
    ds_us, dz_us = _grad(us, pw, ax, b)
    ds_uz, dz_uz = _grad(uz, pw, ax, b)
    want_s = [us, up, uz, phi + ds_us + dz_uz, dz_us - ds_uz]
    pwf = {k: a.astype(np.float64) for k, a in prob.pw_fluid.items()}
    axf = m.axis_fluid.astype(bool)
    ir = prob.inv_rho_fluid.astype(np.float64)
    gs, gz = _grad(st["chi"].astype(np.float64), pwf, axf, b)
    fus, fuz = gs * ir, gz * ir
    want_f = [fus, 0 * fus, fuz, _over_s(fus, pwf, axf, b) + _grad(fus, pwf, axf, b)[0] + _grad(fuz, pwf, axf, b)[1],
              0 * fus]
    for v in range(5):
        want = np.zeros(x["npoint_plot"])
        for off, nel, w in ((0, nf, want_f[v]), (nf, ns, want_s[v])):
            sub = w[:, ::2, ::2]
            sel = mask[off:off + nel]
            want[mp[off:off + nel][sel]] = sub[sel]
        tol = 1e-6 if v < 3 else 5e-5
        assert rel_l2(snap[v, 0], want) < tol, (v, rel_l2(snap[v, 0], want))


def test_snapshot_cadence_and_files(tmp_path):
    from oracle import oracle
    prob = _problem("mtr", niter=12, snap_it=5)
    O = oracle.make_loop(prob)
    O.run(12)
    snap = O.xdmf_snapshots()
    assert snap.shape[1] == 3                       # iter 0, 5, 10 (dump_stuff, time_evol_wave.F90:1167)
    x = prob.xdmf
    times = [k * 5 * prob.deltat for k in range(3)]
    paths = write_xdmf(str(tmp_path), 0, x, snap, times, monopole=False)
    npnt, nel = x["npoint_plot"], x["nelem_plot"]
    pts = np.fromfile(paths["xdmf_points_0000.dat"], dtype=">f4").reshape(npnt, 2)
    assert np.array_equal(pts, x["points"])
    grid = np.fromfile(paths["xdmf_grid_0000.dat"], dtype=">i4").reshape(nel, 4)
    assert np.array_equal(grid, x["grid"]) and grid.max() == npnt - 1
    for v, n in enumerate(["s", "p", "z", "trace", "curlip"]):
        rec = np.fromfile(paths[f"xdmf_snap_{n}_0000.dat"], dtype=">f4").reshape(3, npnt)
        assert np.array_equal(rec, snap[v])
    root = ET.parse(paths["xml"]).getroot()         # well-formed, and the structure of formats 733 / 735
    dom = root.find("Domain")
    assert [d.get("Name") for d in dom.findall("DataItem")] == ["grid", "points"]
    coll = dom.find("Grid")
    assert coll.get("CollectionType") == "Temporal"
    grids = coll.findall("Grid")
    assert [g.get("Name") for g in grids] == ["0001", "0002", "0003"]
    g = grids[1]
    assert abs(float(g.find("Time").get("Value")) - times[1]) < 0.01
    assert int(g.find("Topology").get("NumberOfElements")) == nel
    assert [a.get("Name") for a in g.findall("Attribute")] == ["u_s", "u_p", "u_z", "abs", "straintrace", "curlinplane"]
    slab = g.findall("Attribute")[0].find("DataItem")
    start = slab.findall("DataItem")[0].text.split()
    assert start == ["1", "0", "1", "1", "1", str(npnt)]               # snapshot isnap - 1, all points
    assert slab.findall("DataItem")[1].get("Dimensions").split() == ["3", str(npnt)]
    assert slab.findall("DataItem")[1].text.strip() == "xdmf_snap_s_0000.dat"
    ET.parse(paths["meshonly"])
    # monopole: no u_p attribute, no p file
    pm = _problem("explosion", niter=6, snap_it=5)
    Om = oracle.make_loop(pm)
    Om.run(6)
    pmono = write_xdmf(str(tmp_path / "mono"), 3, pm.xdmf, Om.xdmf_snapshots(), [0.0, 5 * pm.deltat], monopole=True)
    assert "xdmf_snap_p_0003.dat" not in pmono
    names = [a.get("Name") for a in ET.parse(pmono["xml"]).getroot().find("Domain").find("Grid").find("Grid").findall("Attribute")]
    assert names == ["u_s", "u_z", "abs", "straintrace", "curlinplane"]


@pytest.mark.gpu
@pytest.mark.parametrize("src", SRCS)
@pytest.mark.parametrize("strict", [True, False])
def test_cuda_xdmf_snapshots_equal_the_oracle(src, strict):
    from axisem_b200 import solver
    from oracle import oracle
    prob = _problem(src, niter=23, snap_it=4, anel=True, dump=True, strain_it=6)
    G = solver.time_loop(prob, strict=strict)
    O = oracle.make_loop(prob)
    st = seeded_state(O, scale=1e-3, fields=("disp", "velo", "chi", "dchi"))
    for L in (G, O):
        apply_state(L, st)
        L.run(9)
        L.run(14)
    g, o = G.xdmf_snapshots(), O.xdmf_snapshots()
    assert g.shape == o.shape and o.shape[1] == 6
    if strict:
        assert np.array_equal(g, o)
        assert np.array_equal(G.snapshots(), O.snapshots())       # the kwf dumps run next to them
    else:
        for v in range(5):
            if o[v].any():
                assert rel_l2(g[v], o[v]) <= 1e-5, v
    assert np.array_equal(G.xdmf_snapshots(2, 3), g[:, 2:5])


def test_native_host_hands_the_snapshots_over(tmp_path):
    """The C++ host (compiled against the oracle's implementation of the header): module variables
    data_io%dump_xdmf, i_arr_xdmf, ... -> axb_set_xdmf -> PREFIX.rankNNNN.xdmf.f32 -> the reference's
    files."""
    import os
    import subprocess
    from axisem_b200.host.problem_bin import save_problem_bin
    from oracle import oracle
    prob = _problem("mtp", niter=11, snap_it=5)
    save_problem_bin(prob, str(tmp_path / "r0.axbp"))
    r = subprocess.run([oracle.build_host(), "--out", str(tmp_path / "out"), "--quiet", str(tmp_path / "r0.axbp")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    O = oracle.make_loop(prob)
    O.run(11)
    want = O.xdmf_snapshots()
    got = np.fromfile(tmp_path / "out.rank0000.xdmf.f32", dtype=np.float32).reshape(want.shape)
    assert want.shape[1] == 3 and np.array_equal(got, want)
    paths = write_xdmf(str(tmp_path / "Data"), 0, prob.xdmf, got, [0.0, 5 * prob.deltat, 10 * prob.deltat], monopole=False)
    assert os.path.getsize(paths["xdmf_snap_trace_0000.dat"]) == 4 * 3 * prob.xdmf["npoint_plot"]
