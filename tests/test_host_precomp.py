"""The native pre-computation (axisem_b200/hostcxx/precomp.cpp + mapping.cpp): from the MESHER's
databases to the complete time-loop inputs with no Python in between — SURVEY.md section 8(f)
item 1, second half (def_precomp_terms, get_model, analytic_mapping, attenuation set-up, source,
receivers).

  * every array it produces against the Python builder's (axisem_b200/host/precomp.py, the
    restatement the parity tests rest on): to 1 ulp of real(4) wherever the value is significant,
    noise-level otherwise (planes that cancel analytically for concentric elements);
  * the reference's own self-checks: mass = volume (def_grid.f90:1188) and the S/F boundary
    term = 2 per boundary (def_precomp_terms.f90:2743);
  * the element mappings of all four element types against finite differences and their defining
    geometry, and a database whose inner layer is re-typed linear / semino / semiso;
  * the solver started from databases alone gives the seismograms of the Python-built problem.
"""
import os
import subprocess

import numpy as np
import pytest

from axisem_b200.host import AttenuationModel, SourceParams, build_problem, prem_mesh_spec
from axisem_b200.host.meshdb_io import read_axbprob, write_meshdb
from axisem_b200.host.problem_bin import problem_records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_precomp")
COLAT = [10.0, 47.0, 93.0, 131.0, 170.0]


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["bash", os.path.join(ROOT, "axisem_b200", "hostcxx", "build.sh")])
    return EXE


def _run(exe, tmp, probs, src, anel, ani, extra=()):
    files = []
    for r, p in enumerate(probs):
        f = os.path.join(tmp, f"meshdb.dat{r:04d}")
        write_meshdb(p.mesh, f, dt=p.deltat, bkgrdmodel="prem_ani" if ani else "prem_iso")
        files.append(f)
    cmd = [exe, "--out", os.path.join(tmp, "pre"), "--src", src, "--period", "40", "--niter", str(probs[0].niter),
           "--strain-it", "10", "--energy", "--receivers", ",".join(map(str, COLAT))] + list(extra)
    if anel != "none":
        cmd += ["--attenuation", anel]
    out = subprocess.run(cmd + files, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    checks = dict(line.split() for line in out.stdout.strip().splitlines())
    return [read_axbprob(os.path.join(tmp, f"pre.rank{r:04d}.axbp")) for r in range(len(probs))], checks, files


@pytest.mark.parametrize("src,ani,anel,nranks,nranks_r", [("explosion", False, "none", 1, 1), ("mtr", False, "none", 2, 1),
                                                          ("mtp", True, "cg4", 2, 1), ("mtr", True, "full", 4, 1),
                                                          ("explosion", True, "cg4", 2, 1), ("mtr", True, "cg4", 4, 2),
                                                          ("mtp", False, "cg4", 6, 3)])
def test_native_precomputation_equals_the_python_builder(exe, tmp_path, src, ani, anel, nranks, nranks_r):
    spec = prem_mesh_spec(ntheta=16, nr_target=18, anisotropic=ani)
    att = AttenuationModel(coarse_grained=(anel == "cg4")) if anel != "none" else None
    probs = [build_problem(spec, SourceParams(src_type2=src, t_0=40.0), anel=anel != "none", att=att, niter=30, rank=r,
                           nranks=nranks, nranks_r=nranks_r, rec_colat_deg=COLAT, dump=True, strain_it=10, energy=True)
             for r in range(nranks)]
    got, checks, _ = _run(exe, str(tmp_path), probs, src, anel, ani)
    # the reference's self-checks
    assert abs(float(checks["mass_over_volume"]) - 1.0) < 1e-9          # mass (rho = 1) = volume of the shell
    assert int(checks["n_sf_boundaries"]) == 2
    assert abs(float(checks["bdry_sum"]) - 2.0 * 2) < 1e-9              # int_0^pi sin = 2 per boundary
    nrec = 0
    for p, g in zip(probs, got):
        recs = {n: np.ascontiguousarray(a, dtype=d) for n, a, d in problem_records(p)}
        # scale of a family of planes: terms that cancel analytically are compared against it
        fam = {}
        for n, a in recs.items():
            if a.dtype == np.float32 and a.size:
                key = (n.split("%")[0], a.shape[-2:] if a.ndim >= 2 else ())
                fam[key] = max(fam.get(key, 0.0), float(np.abs(a).max()))
        for n, a in recs.items():
            if n not in g:
                assert a.size == 0, n          # a block without fluid (or solid) elements
                continue
            b = g[n]
            assert a.size == b.size, (n, a.shape, b.shape)
            a, b = a.reshape(-1), b.reshape(-1)
            if a.dtype == np.int32:
                assert np.array_equal(a, b), n
            elif a.dtype == np.float64:
                assert np.allclose(a, b, rtol=1e-13, atol=0), n
            else:
                key = (n.split("%")[0], recs[n].shape[-2:] if recs[n].ndim >= 2 else ())
                scale = fam[key]
                big = np.abs(a) > 1e-6 * scale
                ulp = np.abs(a.view(np.int32).astype(np.int64) - b.astype(np.float32).view(np.int32).astype(np.int64))
                assert (ulp[big] <= 1).all(), (n, int(ulp[big].max()))
                assert np.abs(a - b)[~big].max(initial=0.0) <= 1e-12 * scale, n
        nrec += int(g["data_mesh%num_rec"])
    assert nrec == len(COLAT)


def test_element_mappings(exe, tmp_path):
    """curved / linear / semino / semiso through the tool's --mapping-check: corners hit the
    control nodes, derivatives agree with central differences, the straight side of a semi element
    is straight and its curved side lies on the ellipse through its end nodes."""
    out = subprocess.run([exe, "--mapping-check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    vals = dict(line.split()[:2] for line in out.stdout.strip().splitlines())
    for k in ("curved", "linear", "semino", "semiso"):
        assert float(vals[k + "_corner_err"]) < 1e-6, (k, vals)              # metres
        assert float(vals[k + "_derivative_err"]) < 1e-6, (k, vals)          # relative
    assert float(vals["semino_line_err"]) < 1e-6 and float(vals["semiso_line_err"]) < 1e-6
    assert float(vals["semino_ellipse_err"]) < 1e-9 and float(vals["semiso_ellipse_err"]) < 1e-9


def test_database_with_other_element_types(exe, tmp_path):
    """The innermost solid layer re-typed (linear: 8-node serendipity through the true arc mid-points;
    semiso: elliptic bottom, straight top; semino on the row above: straight bottom, elliptic top),
    as the mesher does around its inner cube: the reader takes the types from the database, the
    mappings are dispatched per element, and mass = volume still holds to the accuracy the straight
    sides and the quadrature allow (the two straight-sided rows cut the same lens in and out)."""
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    prob = build_problem(spec, SourceParams(src_type2="explosion", t_0=40.0), niter=10, rec_colat_deg=COLAT)
    f = str(tmp_path / "meshdb.dat0000")
    m = prob.mesh
    ir0 = int(m.solid.ir.min())
    eltype = np.array([b"curved"] * (m.nel_solid + m.nel_fluid))
    eltype[:m.nel_solid][m.solid.ir == ir0] = b"linear"
    write_meshdb(m, f, dt=prob.deltat, eltype=eltype)
    out = subprocess.run([exe, "--out", str(tmp_path / "lin"), "--niter", "10", f], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    c = dict(line.split() for line in out.stdout.strip().splitlines())
    # the parabola through three points of an arc of 11 degrees leaves a relative area error of ~1e-6
    assert abs(float(c["mass_over_volume"]) - 1.0) < 1e-5
    g = read_axbprob(str(tmp_path / "lin.rank0000.axbp"))
    ref = {n: np.asarray(a) for n, a, d in problem_records(prob)}
    lin = (m.solid.ir == ir0)
    a, b = ref["data_matr%M21s"].reshape(-1, 25), g["data_matr%M21s"].reshape(-1, 25)
    assert np.array_equal(a[~lin], b[~lin])                        # untouched elements: identical
    assert not np.array_equal(a[lin], b[lin]) and np.allclose(a[lin], b[lin], rtol=2e-2)
    # semiso below semino: straight interface between two rows
    eltype[:] = b"curved"
    eltype[:m.nel_solid][m.solid.ir == ir0] = b"semiso"
    eltype[:m.nel_solid][m.solid.ir == ir0 + 1] = b"semino"
    write_meshdb(m, f, dt=prob.deltat, eltype=eltype)
    out = subprocess.run([exe, "--out", str(tmp_path / "semi"), "--niter", "10", f], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    c = dict(line.split() for line in out.stdout.strip().splitlines())
    # what semiso loses semino gains; what remains is the GLL quadrature of the two distorted rows
    assert abs(float(c["mass_over_volume"]) - 1.0) < 2e-5


def test_solver_from_databases_alone(exe, tmp_path):
    """axisem_b200_solver (here: its CPU twin linked against the oracle) started from the mesher's
    databases only gives the seismograms of the run set up through Python."""
    from axisem_b200.capi import connect_local, run_group
    from oracle import oracle
    nranks, n = 2, 60
    spec = prem_mesh_spec(ntheta=16, nr_target=18, anisotropic=True)
    probs = [build_problem(spec, SourceParams(src_type2="mtr", t_0=3.0), anel=True, niter=n, rank=r, nranks=nranks,
                           rec_colat_deg=COLAT) for r in range(nranks)]
    files = []
    for r, p in enumerate(probs):
        f = str(tmp_path / f"meshdb.dat{r:04d}")
        write_meshdb(p.mesh, f, dt=p.deltat, bkgrdmodel="prem_ani")
        files.append(f)
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / "run"), "--src", "mtr", "--period", "3",
                          "--niter", str(n), "--attenuation", "cg4", "--receivers", ",".join(map(str, COLAT))] + files,
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    lib = oracle.load()
    loops = [oracle.make_loop(p) for p in probs]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    seen = 0
    for r, (p, L) in enumerate(zip(probs, loops)):
        if not p.num_rec:
            continue
        raw = np.fromfile(tmp_path / f"run.rank{r:04d}.seis.f32", dtype=np.float32).reshape(-1, p.num_rec, 3)
        ref = L.seismograms()
        assert raw.shape == ref.shape and np.abs(ref).max() > 0
        assert np.abs(raw - ref).max() <= 2e-6 * np.abs(ref).max()
        seen += p.num_rec
    assert seen == len(COLAT)


def test_native_xdmf_plot_maps_equal_the_python_builder(exe, tmp_path):
    """dump_xdmf_grid (meshes_io.F90:110-437) in the native pre-computation: masks, mapping, plot
    points and the quadrilateral grid, on two slices with a restricted plot region; and the solver
    started from the databases writes the snapshots of the run set up through Python."""
    from oracle import oracle
    nranks, n = 2, 12
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    opts = dict(rmin=3.0e6, rmax=6.2e6, thetamin=0.0, thetamax=2.0)
    probs = [build_problem(spec, SourceParams(src_type2="mtr", t_0=40.0), niter=n, rank=r, nranks=nranks,
                           rec_colat_deg=COLAT, snap_it=5, xdmf_opts=opts) for r in range(nranks)]
    got, _, files = _run(exe, str(tmp_path), probs, "mtr", "none", False,
                         extra=["--snap-it", "5", "--xdmf-region", "3000", "6200", "0", str(np.degrees(2.0))])
    for p, g in zip(probs, got):
        x = p.xdmf
        assert int(g["data_io%dump_xdmf"]) == 1 and int(g["data_time%snap_it"]) == 5
        assert int(g["data_mesh%npoint_plot"]) == x["npoint_plot"] and int(g["data_mesh%nelem_plot"]) == x["nelem_plot"]
        assert np.array_equal(g["data_mesh%plotting_mask"].reshape(x["plotting_mask"].shape), x["plotting_mask"])
        assert np.array_equal(g["data_mesh%mapping_ijel_iplot"].reshape(x["mapping_ijel_iplot"].shape), x["mapping_ijel_iplot"])
        assert np.array_equal(g["data_mesh%xdmf_grid"].reshape(-1, 4), x["grid"])
        pts = g["data_mesh%xdmf_points"].reshape(-1, 2)
        assert np.abs(pts - x["points"]).max() <= 1e-6 * 6.4e6
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / "run"), "--rundir", str(tmp_path / "RUN"),
                          "--src", "mtr", "--period", "40",
                          "--niter", str(n), "--snap-it", "5", "--receivers", ",".join(map(str, COLAT))] + files,
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    full = [build_problem(spec, SourceParams(src_type2="mtr", t_0=40.0), niter=n, rank=r, nranks=nranks,
                          rec_colat_deg=COLAT, snap_it=5) for r in range(nranks)]
    from axisem_b200.capi import connect_local, run_group
    lib = oracle.load()
    loops = [oracle.make_loop(p) for p in full]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    for r, L in enumerate(loops):
        want = L.xdmf_snapshots()
        raw = np.fromfile(tmp_path / f"run.rank{r:04d}.xdmf.f32", dtype=np.float32).reshape(want.shape)
        assert want.shape[1] == 3
        scale = np.abs(want).max(axis=(1, 2), keepdims=True) + 1e-30
        assert (np.abs(raw - want) / scale).max() <= 1e-5
        # --rundir: the native host writes the reference's Data/xdmf_* files itself — byte for byte what the
        # Python writer (tests/test_xdmf.py holds it against the reference's formats) makes of the same arrays
        from axisem_b200.host.xdmf import write_xdmf
        x = full[r].xdmf                       # (this run plots the whole domain)
        app = f"{r:04d}"
        pts = np.fromfile(tmp_path / "RUN" / "Data" / f"xdmf_points_{app}.dat", dtype=">f4").reshape(-1, 2)
        grd = np.fromfile(tmp_path / "RUN" / "Data" / f"xdmf_grid_{app}.dat", dtype=">i4").reshape(-1, 4)
        assert pts.shape == x["points"].shape and np.abs(pts - x["points"]).max() <= 1e-6 * 6.4e6 and np.array_equal(grd, x["grid"])
        maps = dict(npoint_plot=x["npoint_plot"], nelem_plot=x["nelem_plot"], points=pts, grid=grd)
        ref = write_xdmf(str(tmp_path / f"py{r}"), r, maps, raw, [k * 5 * full[r].deltat for k in range(3)], monopole=False)
        assert len(ref) == 9
        for path in ref.values():
            name = os.path.basename(path)
            assert open(path, "rb").read() == open(tmp_path / "RUN" / "Data" / name, "rb").read(), name
    info = open(tmp_path / "RUN" / "simulation.info").read().splitlines()
    assert int(info[17].split()[0]) == n // 5 and float(info[18].split()[0]) == pytest.approx(5 * full[0].deltat, abs=1e-6)


@pytest.mark.parametrize("stf,choice", [("gauss_1", "gaussi"), ("gauss_2", "gaussi"), ("errorf", "gaussi"),
                                        ("dirac_0", "gaussi"), ("dirac_0", "1dirac"), ("dirac_1", "triang"),
                                        ("dirac_0", "cauchy"), ("dirac_0", "caulor"), ("dirac_0", "sincfc"),
                                        ("quheavi", "gaussi"), ("quheavi", "1dirac")])
def test_native_source_time_functions(exe, tmp_path, stf, choice):
    """compute_stf of the C++ host (every SOURCE_FUNCTION of the reference, every discrete Dirac of
    delta_src) equals the numpy restatement sample by sample; the stf_type code and the shift that
    the symplectic loop would use travel with it."""
    from axisem_b200.capi import STF_TYPES
    from axisem_b200.host.source import stf_shift
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    shift = 7.0 if stf in ("dirac_0", "dirac_1", "quheavi") else None
    src = SourceParams(src_type2="explosion", t_0=4.0 if shift else 40.0, stf_type=stf, discrete_choice=choice, shift_seconds=shift)
    prob = build_problem(spec, src, niter=300, rec_colat_deg=COLAT, dump=True, strain_it=10, energy=True)
    extra = ["--stf", stf, "--discrete-choice", choice, "--period", str(src.t_0), "--niter", "300"]
    if shift:
        extra += ["--shift", str(shift)]
    got, _, _ = _run(exe, str(tmp_path), [prob], "explosion", "none", False, extra)
    a, b = prob.stf, got[0]["data_source%stf"].astype(np.float32).reshape(-1)
    assert a.shape == b.shape and np.abs(a).max() > 0
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    big = np.abs(a) > 1e-6 * np.abs(a).max()
    assert (ulp[big] <= 1).all(), int(ulp[big].max())
    assert np.abs(a - b)[~big].max(initial=0.0) <= 1e-12 * np.abs(a).max()
    assert int(got[0]["data_source%stf_type"]) == STF_TYPES[stf]
    assert float(got[0]["data_source%shift_fact"]) == stf_shift(src, prob.deltat)


def test_external_background_model(exe, tmp_path):
    """bkgrdmodel = 'external' (--ext-model FILE.bm): the reference's own tabulation of prem_ani
    (TESTING/TEST04 model.bm, tests/golden/prem_ani_model_bm.npz) through read_ext_model / get_ext_disc
    / interpolate gives the planes of the analytic prem_ani to the resolution of the table, with
    attenuation (Q from the table's columns)."""
    from .test_background_models import ANI_COLS, _fixture, _write_bm
    spec = prem_mesh_spec(ntheta=16, nr_target=18, anisotropic=True)
    att = AttenuationModel(coarse_grained=True)
    prob = build_problem(spec, SourceParams(src_type2="mtr", t_0=40.0), anel=True, att=att, niter=30,
                         rec_colat_deg=COLAT, dump=True, strain_it=10, energy=True)
    ana, _, _ = _run(exe, str(tmp_path), [prob], "mtr", "cg4", True)
    bm = str(tmp_path / "prem_ani.bm")
    _write_bm(bm, _fixture()[0], ANI_COLS)
    ext, checks, _ = _run(exe, str(tmp_path), [prob], "mtr", "cg4", True, ["--ext-model", bm])
    assert abs(float(checks["mass_over_volume"]) - 1.0) < 1e-9
    n = 0
    for name, a in ana[0].items():
        b = ext[0][name]
        assert a.shape == b.shape, name
        if a.dtype.kind in "iu" or not a.size:
            assert np.array_equal(a, b), name
            continue
        a64, b64 = a.astype(np.float64).reshape(-1), b.astype(np.float64).reshape(-1)
        scale = np.abs(a64).max()
        if name.startswith(("data_matr%", "attenuation%")) and scale > 0:
            assert np.abs(a64 - b64).max() <= 1e-3 * scale, (name, np.abs(a64 - b64).max() / scale)
            n += 1
        elif name.startswith(("data_mesh%", "data_spec%", "data_pointwise%", "data_source%", "data_comm%")):
            if name != "data_mesh%bkgrdmodel":
                assert np.allclose(a64, b64, rtol=1e-3, atol=1e-3 * scale), name
    assert n > 30
    # the planes do differ: the table is linear between its nodes
    m = "data_matr%M11s"
    assert not np.array_equal(ana[0][m], ext[0][m])
    assert np.array_equal(ana[0]["data_matr%Q_mu"], ext[0]["data_matr%Q_mu"])     # piecewise constant: exact


def test_receiver_files_and_rotation(exe, tmp_path):
    """prepare_from_recfile_seis on the native host: receivers.dat (colatlon) and STATIONS (stations,
    repeated station + network dropped, longitudes <= 0 shifted), the rotation of rotations.f90 into
    the frame with the source on the pole (epicentral distance = colatitude there; the post-processing's
    receiver_location, restated from another file of the reference, maps it back), receiver_pts.dat and
    the location of the closest surface points."""
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    probs = [build_problem(spec, SourceParams(src_type2="explosion", t_0=40.0), niter=10, rank=r, nranks=2,
                           rec_colat_deg=COLAT, dump=True, strain_it=10, energy=True) for r in range(2)]
    st = tmp_path / "STATIONS"
    st.write_text("AAK II 42.639 74.494 1645.0 30.0\n"
                  "ANMO IU 34.946 -106.457 1850.0 100.0\n"
                  "AAK II 42.639 74.494 1645.0 30.0\n"          # the same station again: dropped
                  "SPA IU -89.93 145.0 2927.0 0.0\n"
                  "AAK KN 42.639 74.494 1645.0 30.0\n")         # same name, another network: kept
    src_lat, src_lon = 36.5, 140.25
    got, _, _ = _run(exe, str(tmp_path), probs, "explosion", "none", False,
                     ["--stations", str(st), "--src-lat", str(src_lat), "--src-lon", str(src_lon)])
    pre = str(tmp_path / "pre")
    names = [l.split() for l in open(pre + ".receiver_names.dat").read().strip().splitlines()]
    assert [n[0] for n in names] == ["AAK_II", "ANMO_IU", "SPA_IU", "AAK_KN"]
    colat = np.array([float(n[1]) for n in names])
    lon = np.array([float(n[2]) for n in names])
    assert np.allclose(colat, [90 - 42.639, 90 - 34.946, 179.93, 90 - 42.639]) and np.allclose(lon, [74.494, 360 - 106.457, 145.0, 74.494])
    rot = np.loadtxt(pre + ".receiver_rotated.dat")
    # epicentral distance by the spherical law of cosines
    sc, sl = np.radians(90 - src_lat), np.radians(src_lon)
    rc, rl = np.radians(colat), np.radians(lon)
    dist = np.degrees(np.arccos(np.cos(sc) * np.cos(rc) + np.sin(sc) * np.sin(rc) * np.cos(rl - sl)))
    assert np.allclose(rot[:, 0], dist, atol=1e-6)
    # back through the inverse rotation: x = R y with R = rot_mat of def_rot_matrix
    R = np.array([[np.cos(sc) * np.cos(sl), -np.sin(sl), np.sin(sc) * np.cos(sl)],
                  [np.cos(sc) * np.sin(sl), np.cos(sl), np.sin(sc) * np.sin(sl)],
                  [-np.sin(sc), 0.0, np.cos(sc)]])
    th, ph = np.radians(rot[:, 0]), np.radians(rot[:, 1])
    x = R @ np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
    x0 = np.stack([np.sin(rc) * np.cos(rl), np.sin(rc) * np.sin(rl), np.cos(rc)])
    assert np.abs(x - x0).max() < 1e-6          # (the reference's + smallval_dble costs the longitude of a polar station 0.005 deg)
    assert np.allclose((np.degrees(np.arctan2(x[1], x[0])) % 360.0)[[0, 1, 3]], lon[[0, 1, 3]], atol=1e-5)
    # the grid points taken: each receiver on exactly one rank, within half a surface element of its distance
    pts = np.loadtxt(pre + ".receiver_pts.dat")
    assert pts.shape == (4, 3) and np.allclose(pts[:, 1], rot[:, 1])
    assert np.abs(pts[:, 0] - rot[:, 0]).max() < 180.0 / 16 / 2
    nrec = 0
    for r, g in enumerate(got):
        idx = g["data_mesh%loc2globrec"].reshape(-1) if int(g["data_mesh%num_rec"]) else np.zeros(0, int)
        assert (pts[idx - 1, 2] == r).all()
        assert np.allclose(g["data_mesh%recfile_th"].reshape(-1), pts[idx - 1, 0]) if idx.size else True
        nrec += idx.size
    assert nrec == 4
    # receivers.dat: source on the pole (no rotation), names recfile_NNNN
    rd = tmp_path / "receivers.dat"
    rd.write_text("3\n10.0 0.0\n95.5 270.0\n171.0 12.5\n")
    _run(exe, str(tmp_path), probs, "explosion", "none", False, ["--receivers-file", str(rd)])
    names = [l.split() for l in open(pre + ".receiver_names.dat").read().strip().splitlines()]
    assert [n[0] for n in names] == ["recfile_0001", "recfile_0002", "recfile_0003"]
    assert np.allclose(np.loadtxt(pre + ".receiver_rotated.dat"), [[10.0, 0.0], [95.5, 270.0], [171.0, 12.5]])
    # the reference's stops
    rd.write_text("1\n10.0 -5.0\n")
    files = [str(tmp_path / f"meshdb.dat{r:04d}") for r in range(2)]
    out = subprocess.run([exe, "--out", pre, "--receivers-file", str(rd)] + files, capture_output=True, text=True)
    assert out.returncode != 0 and "negative receiver longitudes" in out.stderr


def _fortran_field(line, width):
    return line[:width], line[width:]


@pytest.mark.parametrize("src,nranks", [("mtr", 2), ("explosion", 1)])
def test_solver_leaves_the_references_run_directory(exe, tmp_path, src, nranks):
    """--rundir: what the reference leaves behind with USE_NETCDF false and its post-processing reads —
    simulation.info in the formats of parameters.F90:1410-1465 (read back the way post_processing.F90:614-647
    reads it), Data/receiver_names.dat, Data/receiver_pts.dat, and one <receiver>_disp.dat per station
    with the samples of the run (two columns for a monopole, list order of the STATIONS file across ranks)."""
    from axisem_b200.capi import connect_local, run_group
    from oracle import oracle
    n = 40
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    st = tmp_path / "STATIONS"
    st.write_text("S170 XX -80.0 0.0 0.0 0.0\nS010 XX 80.0 0.0 0.0 0.0\nS093 YY -3.0 45.0 0.0 0.0\nS047 XX 43.0 200.0 0.0 0.0\n")
    colat = [170.0, 10.0, 93.0, 47.0]
    probs = [build_problem(spec, SourceParams(src_type2=src, t_0=3.0), niter=n, rank=r, nranks=nranks, seis_it=2,
                           rec_colat_deg=colat) for r in range(nranks)]
    files = []
    for r, p in enumerate(probs):
        f = str(tmp_path / f"meshdb.dat{r:04d}")
        write_meshdb(p.mesh, f, dt=p.deltat, bkgrdmodel="prem_iso")
        files.append(f)
    rundir = tmp_path / "RUN"
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / "run"), "--rundir", str(rundir), "--src", src,
                          "--period", "3", "--niter", str(n), "--seis-it", "2", "--stations", str(st)] + files,
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    # ---- simulation.info: 32 lines, value in a 20 / 22 / 15 wide field, label right-justified in 45
    lines = open(rundir / "simulation.info").read().splitlines()
    assert len(lines) == 32
    widths = [20, 22, 20, 20, 20, 20, 20, 22, 22, 22, 22, 15, 20, 20, 22, 20, 22, 20, 22, 20, 20, 20, 22, 20, 20, 20, 20, 22, 20, 20, 20, 20]
    labels = ["background model", "time step [s]", "number of time steps", "source type", "source type", "source time function", "simtype",
              "dominant source period", "source depth [km]", "Source colatitude", "Source longitude", "scalar source magnitude",
              "number of receivers", "length of seismogram [time samples]", "seismogram sampling [s]", "number of strain dumps",
              "strain dump sampling rate [s]", "number of snapshot dumps", "snapshot dump sampling rate [s]", "receiver components",
              "ibeg: beginning gll index for wavefield dum",       # 47 characters through an a45 edit descriptor
              "iend: end gll index for wavefield dumps", "source shift factor [s]",
              "source shift factor for deltat", "source shift factor for seis_dt", "source shift factor for deltat_coarse",
              "receiver file type", "receiver spacing (0 if not even)", "use netcdf for wavefield output?", "nelem", "nel_fluid", "nproc"]
    vals = []
    for line, w, lab in zip(lines, widths, labels):
        assert len(line) == w + 45, line
        v, l = _fortran_field(line, w)
        assert l.strip() == lab, line
        vals.append(v.split()[0])
    p0 = probs[0]
    assert vals[0] == "prem_iso" and float(vals[1]) == pytest.approx(p0.deltat, abs=1e-7) and int(vals[2]) == n
    assert vals[3] == ("dipole" if src == "mtr" else "monopole") and vals[4] == src and vals[5] == "gauss_0" and vals[6] == "single"
    assert float(vals[7]) == 3.0 and float(vals[8]) == 100.0 and float(vals[9]) == 0.0 and float(vals[10]) == 0.0
    assert vals[11] == "1.00000E+20" and int(vals[12]) == 4 and int(vals[13]) == n // 2 + 1
    assert float(vals[14]) == pytest.approx(2 * p0.deltat, abs=1e-6)
    assert vals[19] == "cyl" and vals[26] == "stations" and vals[28] == "F" and int(vals[31]) == nranks
    assert int(vals[29]) == p0.mesh.nel_solid + p0.mesh.nel_fluid and int(vals[30]) == p0.mesh.nel_fluid
    shift = float(vals[22])
    assert int(vals[23]) == round(shift / p0.deltat) and int(vals[24]) == round(shift / (2 * p0.deltat))
    # ---- receivers
    names = [l.split()[0] for l in open(rundir / "Data" / "receiver_names.dat").read().strip().splitlines()]
    assert names == ["S170_XX", "S010_XX", "S093_YY", "S047_XX"]
    pts = np.loadtxt(rundir / "Data" / "receiver_pts.dat")
    assert pts.shape == (4, 3) and np.abs(pts[:, 0] - colat).max() < 180.0 / 16 / 2
    # ---- seismograms: the run itself, per station
    lib = oracle.load()
    loops = [oracle.make_loop(p) for p in probs]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    ref = np.zeros((n // 2 + 1, 4, 3), np.float32)
    for p, L in zip(probs, loops):
        if p.num_rec:
            ref[:, p.rec_index, :] = L.seismograms()
    assert np.abs(ref).max() > 0
    for k, name in enumerate(names):
        tab = np.loadtxt(rundir / "Data" / f"{name}_disp.dat")
        assert tab.shape == (n // 2 + 1, 2 if src == "explosion" else 3)
        want = ref[:, k, :][:, [0, 2]] if src == "explosion" else ref[:, k, :]
        assert np.abs(tab - want).max() <= 3e-6 * np.abs(ref).max() + 1e-8 * np.abs(want).max()
    stf = np.loadtxt(rundir / "Data" / "stf_seis.dat")
    assert stf.shape == (n // 2, 2) and np.allclose(stf[:, 1], p0.stf[1::2], rtol=1e-6)
    # ---- SAVE_ENERGY: energy_sol / _flu / _glob.dat, one line per time step from t = 0, 1pe16.6 columns
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / "run"), "--rundir", str(rundir), "--src", src,
                          "--period", "3", "--niter", str(n), "--seis-it", "2", "--stations", str(st), "--energy"] + files,
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    eprobs = [build_problem(spec, SourceParams(src_type2=src, t_0=3.0), niter=n, rank=r, nranks=nranks, seis_it=2,
                            rec_colat_deg=colat, energy=True) for r in range(nranks)]
    loops = [oracle.make_loop(p) for p in eprobs]
    connect_local(lib, loops)
    run_group(lib, loops, n)
    e = 2 * np.pi * sum(L.energy().astype(np.float64) for L in loops)          # (n + 1, 4): epot_s, ekin_s, epot_f, ekin_f
    sol, flu, glob = (np.loadtxt(rundir / "Data" / f"energy_{k}.dat") for k in ("sol", "flu", "glob"))
    line = open(rundir / "Data" / "energy_glob.dat").readline().rstrip("\n")
    assert len(line) == 64 and sol.shape == (n + 1, 3) and glob.shape == (n + 1, 4)
    tt = np.arange(n + 1) * p0.deltat
    scale = np.abs(e).max()
    assert np.allclose(sol[:, 0], tt, rtol=2e-6, atol=1e-9) and scale > 0
    assert np.abs(sol[:, 1] - e[:, 1]).max() <= 3e-6 * scale and np.abs(sol[:, 2] - e[:, 0]).max() <= 3e-6 * scale
    assert np.abs(flu[:, 1] - e[:, 2]).max() <= 3e-6 * scale and np.abs(flu[:, 2] - e[:, 3]).max() <= 3e-6 * scale
    assert np.abs(glob[:, 1] - (e[:, 0] + e[:, 2])).max() <= 3e-6 * scale
    assert np.abs(glob[:, 3] - 0.5 * e.sum(axis=1)).max() <= 3e-6 * scale
