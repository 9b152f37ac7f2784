"""The output database layout (axisem_b200/host/nc_layout.py) against the definitions of the
reference's `nc_define_outputfile`, extracted from its Fortran into
tests/golden/nc_schema_reference.json, and the data of a (CPU oracle) run against what went in."""
import json
import os

import numpy as np
import pytest

from axisem_b200.host import SourceParams, build_problem, nc_layout
from tests.util import small_spec

HERE = os.path.dirname(os.path.abspath(__file__))
XTYPE = {"NF90_FLOAT": "f4", "NF90_DOUBLE": "f8", "NF90_INT": "i4", "NF90_CHAR": "S1"}


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(HERE, "golden", "nc_schema_reference.json")))


@pytest.mark.parametrize("dump_type", ["displ_only", "strain_only", "fullfields"])
@pytest.mark.parametrize("monopole", [True, False])
def test_schema_is_the_references(ref, dump_type, monopole):
    sch = nc_layout.schema(nrec=5, nseismo=11, niter=40, dump_wavefields=True, dump_type=dump_type, monopole=monopole,
                           npoints_global=123, nstrain=4, nelem_kwf_global=9, anel=True)
    assert list(sch["groups"]) == ref["groups"]
    lens = {"nseismo": 11, "niter": 40, "3": 3, "40": 40, "nrec": 5, "nstrain": 4, "npoints_global": 123,
            "nelem_kwf_global": 9, "4": 4, "npol+1": 5}
    for name, d in ref["dimensions"].items():
        holder = sch["dimensions"] if d["group"] == "" else sch["groups"][d["group"]]["dimensions"]
        if d["group"] == "Mesh" and dump_type != "displ_only":
            assert name not in holder                       # defined inside `if displ_only` in the reference
            continue
        assert holder[name] == lens[d["len"]], name
    snap_names = ref["snapshot_variables"][dump_type][0 if monopole else 1]
    assert snap_names == nc_layout.SNAP_VARS[(dump_type, monopole)]
    mesh_displ_only = {"midpoint_mesh", "eltype", "axis", "fem_mesh", "sem_mesh", "mp_mesh_S", "mp_mesh_Z", "G0", "G1",
                       "G2", "gll", "glj"}
    seen = 0
    for v in ref["variables"]:
        vars_ = sch["variables"] if v["group"] == "" else sch["groups"][v["group"]]["variables"]
        names = snap_names if v["name"].startswith("trim(") else [v["name"]]
        if v["group"] == "Mesh" and v["name"] in mesh_displ_only and dump_type != "displ_only":
            assert v["name"] not in vars_
            continue
        for n in names:
            assert n in vars_, (v["group"], n)
            assert vars_[n]["dtype"] == XTYPE[v["xtype"]], n
            assert vars_[n]["dims"] == v["dims_fortran"][::-1], n      # netCDF order is the reverse of the Fortran dimids
            seen += 1
    total = sum(len(g["variables"]) for g in sch["groups"].values()) + len(sch["variables"])
    assert seen == total                                     # nothing in the layout that the reference does not define


def test_database_of_a_two_rank_run(ref, tmp_path):
    """Two theta-slices through the oracle with a displ_only dump: the directory holds every
    variable of the schema, displacement(receivers, components, time) is the recdumpvar of the
    ranks, the snapshot variables are the ranks' oneddumpvar blocks side by side, sem_mesh indexes
    mesh_S / mesh_Z consistently, and every global attribute of the reference is there."""
    from axisem_b200.capi import connect_local, run_group
    from oracle import oracle
    n, colat = 24, np.linspace(10, 170, 9)
    probs = [build_problem(small_spec(), SourceParams(src_type2="mtr", t_0=4.0), niter=n, dump=True, strain_it=6,
                           seis_it=2, rank=r, nranks=2, rec_colat_deg=colat, anel=True) for r in range(2)]
    lib = oracle.load()
    loops = [oracle.make_loop(p) for p in probs]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    seis = [L.seismograms() for L in loops]
    snaps = [L.snapshots() for L in loops]
    out = str(tmp_path / "run.ncdir")
    sch = nc_layout.write_database(out, probs, seis, snaps, colat_deg=colat)
    for gname, g in sch["groups"].items():
        for v in g["variables"]:
            assert os.path.exists(os.path.join(out, gname, v + ".bin")), (gname, v)
    disp = nc_layout.read_variable(out, "Seismograms", "displacement")
    assert disp.shape == (9, 3, n // 2 + 1)
    for p, s in zip(probs, seis):
        assert np.array_equal(disp[p.rec_index], s.transpose(1, 2, 0))
    assert np.abs(disp).max() > 0
    ds = nc_layout.read_variable(out, "Snapshots", "disp_s")
    np0 = snaps[0].shape[2]
    assert ds.shape == (n // 6 + 1, np0 + snaps[1].shape[2])
    assert np.array_equal(ds[:, :np0], snaps[0][0]) and np.array_equal(ds[:, np0:], snaps[1][0])
    # mesh: the GLL points of an element, through sem_mesh, are where its geometry says they are
    S, Z = (nc_layout.read_variable(out, "Mesh", k) for k in ("mesh_S", "mesh_Z"))
    sem = nc_layout.read_variable(out, "Mesh", "sem_mesh")
    fem = nc_layout.read_variable(out, "Mesh", "fem_mesh")
    mid = nc_layout.read_variable(out, "Mesh", "midpoint_mesh")
    from axisem_b200.host.mesh import element_coords
    m0 = probs[0].mesh
    _, _, _, s_el, z_el, *_ = element_coords(m0.solid, m0.basis)
    assert np.allclose(S[sem[:m0.nel_solid]], s_el) and np.allclose(Z[sem[:m0.nel_solid]], z_el)
    assert np.array_equal(fem[:, 0], sem[:, 0, 0]) and np.array_equal(fem[:, 2], sem[:, 4, 4])
    assert np.array_equal(mid, sem[:, 2, 2])
    assert np.allclose(nc_layout.read_variable(out, "Mesh", "mp_mesh_S"), S[mid])
    r = np.hypot(S, Z)
    vp = nc_layout.read_variable(out, "Mesh", "mesh_vp")
    assert r.max() == pytest.approx(6371e3) and 1e3 < vp.min() and vp.max() < 14e3
    have = set(json.load(open(os.path.join(out, "schema.json")))["attributes"])
    want = {n for n, _ in ref["global_attributes"]}
    assert want <= have, want - have


def test_database_of_a_native_run_equals_the_python_one(tmp_path):
    """The same two-rank run set up twice — by the Python builder, and by the native chain (mesher-format
    databases -> axisem_b200_precomp -> C++ host) — gives the same output database: every variable of every
    group (coordinates, model, element tables, seismograms, snapshots) and the global attributes."""
    import subprocess
    from axisem_b200.capi import connect_local, run_group
    from axisem_b200.host.meshdb_io import write_meshdb
    from oracle import oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n, colat = 24, np.linspace(10, 170, 9)
    probs = [build_problem(small_spec(), SourceParams(src_type2="mtr", t_0=4.0), niter=n, dump=True, strain_it=6,
                           seis_it=2, rank=r, nranks=2, rec_colat_deg=colat, anel=True) for r in range(2)]
    lib = oracle.load()
    loops = [oracle.make_loop(p) for p in probs]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    py = str(tmp_path / "py.ncdir")
    nc_layout.write_database(py, probs, [L.seismograms() for L in loops], [L.snapshots() for L in loops], colat_deg=colat)
    files = []
    for r, p in enumerate(probs):
        files.append(str(tmp_path / f"meshdb.dat{r:04d}"))
        write_meshdb(p.mesh, files[-1], dt=p.deltat, bkgrdmodel="prem_iso")
    pre = subprocess.run([os.path.join(root, "axisem_b200", "axisem_b200_precomp"), "--out", str(tmp_path / "pre"), "--src", "mtr",
                          "--period", "4", "--niter", str(n), "--seis-it", "2", "--strain-it", "6", "--attenuation", "cg4",
                          "--receivers", ",".join(map(str, colat))] + files, capture_output=True, text=True)
    assert pre.returncode == 0, pre.stderr
    cont = [str(tmp_path / f"pre.rank{r:04d}.axbp") for r in range(2)]
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / "run")] + cont, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    tool = subprocess.run([os.sys.executable, os.path.join(root, "tools", "native_to_nc_layout.py"), "--out", str(tmp_path / "nat.ncdir"),
                           "--run", str(tmp_path / "run")] + cont, capture_output=True, text=True)
    assert tool.returncode == 0, tool.stderr
    nat = str(tmp_path / "nat.ncdir")
    a, b = (json.load(open(os.path.join(d, "schema.json"))) for d in (py, nat))
    assert a["dimensions"] == b["dimensions"] and a["groups"].keys() == b["groups"].keys()
    for k, v in a["attributes"].items():
        w = b["attributes"][k]
        assert (v == pytest.approx(w, rel=1e-6) if isinstance(v, float) else v == w), (k, v, w)
    nvar = 0
    for gname, g in [("", a)] + list(a["groups"].items()):
        for v in g["variables"]:
            x, y = nc_layout.read_variable(py, gname, v), nc_layout.read_variable(nat, gname, v)
            assert x.shape == y.shape, (gname, v)
            if x.dtype.kind in "iuS":
                assert np.array_equal(x, y), (gname, v)
            else:
                scale = np.abs(x).max()
                assert np.abs(x.astype(np.float64) - y).max() <= 3e-6 * scale + 1e-30, (gname, v, np.abs(x.astype(np.float64) - y).max() / (scale + 1e-300))
            nvar += 1
    assert nvar > 40


def test_database_of_a_native_run_on_a_whole_earth(tmp_path):
    """The same tool on a mesh the Python builder cannot make: inner square, ring, fluid core, coarsening
    layer, two ranks (tests/doubling_mesh.py).  The Mesh group carries the four element types, one point per
    global number of either domain from the centre to the surface, and element tables that index it
    consistently; with and without wavefield dumps."""
    import subprocess
    from oracle import oracle
    from axisem_b200.host.spectral import SpectralBasis
    from . import doubling_mesh as dm
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows = [(1221.5e3, 2350e3, "R"), (2350e3, 3480e3, "R"), (3480e3, 3630e3, "R"), (3630e3, 4600e3, "R"), (4600e3, 4900e3, "D"),
            (4900e3, 5600e3, "R"), (5600e3, 5701e3, "R"), (5701e3, 5771e3, "R"), (5771e3, 5971e3, "R"), (5971e3, 6151e3, "R"),
            (6151e3, 6291e3, "R"), (6291e3, 6371e3, "R")]
    F = dm.build_rows(rows, 16, cube_halfwidth=500e3, fluid=lambda r: 1221.5e3 < r < 3480e3)
    parts = dm.partition(F, 2)
    files = []
    for r, P in enumerate(parts):
        files.append(str(tmp_path / f"e.dat{r:04d}"))
        dm.write_database(files[-1], P, SpectralBasis(4), bkgrdmodel="prem_iso_light", solid_domain=[1] * 8 + [0, 1], dt=0.5)
    out = {}
    for tag, extra in (("dump", ["--strain-it", "50"]), ("nodump", [])):
        pre = subprocess.run([os.path.join(root, "axisem_b200", "axisem_b200_precomp"), "--out", str(tmp_path / f"pre_{tag}"),
                              "--src", "explosion", "--depth", "300", "--period", "250", "--niter", "400", "--seis-it", "4",
                              "--receivers", "30,90,150"] + extra + files, capture_output=True, text=True)
        assert pre.returncode == 0, pre.stderr
        cont = [str(tmp_path / f"pre_{tag}.rank{r:04d}.axbp") for r in range(2)]
        run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / f"run_{tag}")] + cont, capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        out[tag] = str(tmp_path / f"{tag}.ncdir")
        sch = nc_layout.write_database_native(out[tag], cont, str(tmp_path / f"run_{tag}"), background_model="prem_iso_light")
        assert sch["attributes"]["background model"] == "prem_iso_light" and sch["attributes"]["source type"] == "explosion"
    d = out["dump"]
    S, Z = (nc_layout.read_variable(d, "Mesh", k) for k in ("mesh_S", "mesh_Z"))
    npt = sum(P["nglob_solid"] + P["nglob_fluid"] for P in parts)
    assert S.size == npt and np.hypot(S, Z).min() == 0.0 and np.hypot(S, Z).max() == pytest.approx(6371e3)
    el = nc_layout.read_variable(d, "Mesh", "eltype")
    assert np.bincount(el).tolist() == [F["eltype"].count(k) for k in ("curved", "linear", "semino", "semiso")]
    sem, fem, mid = (nc_layout.read_variable(d, "Mesh", k) for k in ("sem_mesh", "fem_mesh", "midpoint_mesh"))
    assert sem.shape == (F["nelem"], 5, 5) and sem.min() == 0 and sem.max() == npt - 1
    assert np.array_equal(fem[:, 0], sem[:, 0, 0]) and np.array_equal(fem[:, 2], sem[:, 4, 4]) and np.array_equal(mid, sem[:, 2, 2])
    assert np.allclose(nc_layout.read_variable(d, "Mesh", "mp_mesh_S"), S[mid])
    # the corners of an element, through fem_mesh, are the corners the generator gave it (rank blocks in order)
    corners = np.array([c for P in parts for c in P_corners(P, F)])
    assert np.allclose(np.stack([S[fem], Z[fem]], axis=-1), corners, atol=1e-3)
    vs = nc_layout.read_variable(d, "Mesh", "mesh_vs")
    r = np.hypot(S, Z)
    assert (vs[(r > 1221.5e3 + 1) & (r < 3480e3 - 1)] == 0).all() and (vs[r > 3480e3 + 1] > 3000).all()       # fluid outer core
    ds = nc_layout.read_variable(d, "Snapshots", "disp_s")
    assert ds.shape == (400 // 50 + 1, npt) and np.abs(ds).max() > 0
    disp = nc_layout.read_variable(d, "Seismograms", "displacement")
    assert disp.shape == (3, 3, 101) and np.array_equal(disp, nc_layout.read_variable(out["nodump"], "Seismograms", "displacement"))
    assert not os.path.exists(os.path.join(out["nodump"], "Mesh", "mesh_S.bin"))


def P_corners(P, F):
    """corner coordinates (4, 2) per element of a rank, in the rank's element order"""
    crd = P["crd"].reshape(-1, 8, 2)
    return [crd[e, [0, 2, 4, 6]] for e in range(P["nelem"])]
