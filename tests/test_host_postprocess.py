"""axisem_b200_postproc (hostcxx/postprocess.cpp) against what SOLVER/UTILS/post_processing.F90
does: sum over the four basis runs of a full moment tensor with the azimuthal radiation factors
(:838-918), CMTSOLUTION units (:774-793), rotation of the receiver components for a source
anywhere on the sphere (:187-232, :922-1009), the causal STF convolution (:1014-1084), and the
component order of the output (N E Z / theta phi r).  The expectations are restated here in
numpy, formula by formula."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_postproc")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["bash", os.path.join(ROOT, "axisem_b200", "hostcxx", "build.sh")])
    return EXE


def _rot_mat(sc, sl):
    ct, st, cp, sp = np.cos(sc), np.sin(sc), np.cos(sl), np.sin(sl)
    return np.array([[ct * cp, -sp, st * cp], [ct * sp, cp, st * sp], [-st, 0.0, ct]])


def _prefactor(t, M, mag, lon):
    s = np.asarray(M) / mag
    if t == "mrr":
        return s[0], 0.0, s[0]
    if t == "mtt_p_mpp":
        return s[1] + s[2], 0.0, s[1] + s[2]
    if t in ("mtr", "mpr"):
        a = s[3] * np.cos(lon) + s[4] * np.sin(lon)
        return a, -s[3] * np.sin(lon) + s[4] * np.cos(lon), a
    if t in ("mtp", "mtt_m_mpp"):
        a = (s[1] - s[2]) * np.cos(2 * lon) + 2 * s[5] * np.sin(2 * lon)
        return a, (s[2] - s[1]) * np.sin(2 * lon) + 2 * s[5] * np.cos(2 * lon), a
    raise ValueError(t)


def _expected(runs, M, colat, lon, sc, sl, sys):
    R = _rot_mat(sc, sl)
    nrec = colat.size
    out = []
    for r in range(nrec):
        tot = 0.0
        for t, mag, raw in runs:                           # raw (ns, nrec, 3)
            f = np.array(_prefactor(t, M, mag, lon[r]))
            tot = tot + raw[:, r, :].astype(np.float32) * f[None, :]
            tot = tot.astype(np.float32)
        us, up, uz = tot[:, 0].astype(float), tot[:, 1].astype(float), tot[:, 2].astype(float)
        x0 = np.array([np.sin(colat[r]) * np.cos(lon[r]), np.sin(colat[r]) * np.sin(lon[r]), np.cos(colat[r])])
        x = R @ x0
        x /= np.linalg.norm(x)
        tho = np.arccos(x[2])
        pho = np.arctan2(x[1], x[0]) % (2 * np.pi)
        v = np.stack([np.cos(lon[r]) * us - np.sin(lon[r]) * up, np.sin(lon[r]) * us + np.cos(lon[r]) * up, uz])
        if sc > 0 or sl > 0:
            v = R @ v
        e_r = np.array([np.sin(tho) * np.cos(pho), np.sin(tho) * np.sin(pho), np.cos(tho)])
        e_t = np.array([np.cos(tho) * np.cos(pho), np.cos(tho) * np.sin(pho), -np.sin(tho)])
        e_p = np.array([-np.sin(pho), np.cos(pho), 0.0])
        if sys == "enz":
            o = np.stack([-(e_t @ v), e_p @ v, e_r @ v])
        elif sys == "sph":
            o = np.stack([e_t @ v, e_p @ v, e_r @ v])
        elif sys == "xyz":
            o = v
        else:
            raise ValueError(sys)
        out.append(o)
    return np.array(out)                                   # (nrec, 3, ns)


@pytest.mark.parametrize("sys", ["enz", "sph", "xyz"])
@pytest.mark.parametrize("srcloc", [(0.0, 0.0), (37.5, 143.0)])
def test_four_run_moment_tensor_sum_and_rotation(exe, tmp_path, sys, srcloc):
    rng = np.random.default_rng(5)
    nrec, ns = 7, 64
    colat = np.deg2rad(rng.uniform(5, 175, nrec))
    lon = np.deg2rad(rng.uniform(0, 360, nrec))
    M_dyncm = np.array([1.2e26, -0.7e26, -0.5e26, 2.1e26, -1.4e26, 0.9e26])      # Mrr Mtt Mpp Mrt Mrp Mtp
    cmt = tmp_path / "CMTSOLUTION"
    cmt.write_text(" PDE 2011  3 11  5 46 23.00  38.3200  142.3700  24.4 7.2 9.0 TEST EVENT\n"
                   "event name:     TEST\ntime shift:      0.0000\nhalf duration:   0.0000\n"
                   "latitude:       52.5\nlongitude:     143.0\ndepth:          24.4\n"
                   + "".join(f"{n}:      {v:.6e}\n" for n, v in zip(("Mrr", "Mtt", "Mpp", "Mrt", "Mrp", "Mtp"), M_dyncm)))
    st = tmp_path / "st.txt"
    st.write_text("".join(f"{np.rad2deg(c):.12f} {np.rad2deg(l):.12f}\n" for c, l in zip(colat, lon)))
    runs, args = [], []
    for k, (t, mag) in enumerate((("mrr", 1e20), ("mtt_p_mpp", 2e20), ("mtr", 1e20), ("mtp", 0.5e20))):
        raw = rng.standard_normal((ns, nrec, 3)).astype(np.float32)
        if t in ("mrr", "mtt_p_mpp"):
            raw[:, :, 1] = 0.0
        f = tmp_path / f"run{k}.seis.f32"
        raw.tofile(f)
        runs.append((t, mag, raw))
        args += ["--run", t, repr(mag), str(f)]
    out = tmp_path / "out.f32"
    run = subprocess.run([exe, "--cmt", str(cmt), "--sys", sys, "--srccolat", repr(srcloc[0]), "--srclon", repr(srcloc[1]),
                          "--stations", str(st), "--out", str(out)] + args, capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    got = np.fromfile(out, dtype=np.float32).reshape(nrec, 3, ns)
    want = _expected(runs, M_dyncm / 1e7, colat, lon, np.deg2rad(srcloc[0]), np.deg2rad(srcloc[1]), sys)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-6 * scale


def test_single_run_at_the_pole_has_the_references_component_order(exe, tmp_path):
    """enz -> (N, E, Z); sph -> (theta, phi, r) (post_processing.F90:972-990)."""
    nrec, ns = 3, 16
    rng = np.random.default_rng(2)
    colat = np.deg2rad(np.array([30.0, 90.0, 140.0]))
    lon = np.deg2rad(np.array([0.0, 45.0, 270.0]))
    raw = rng.standard_normal((ns, nrec, 3)).astype(np.float32)
    (tmp_path / "s.f32").write_bytes(raw.tobytes())
    (tmp_path / "st.txt").write_text("".join(f"{np.rad2deg(c):.10f} {np.rad2deg(l):.10f}\n" for c, l in zip(colat, lon)))
    res = {}
    for sys in ("enz", "sph", "cyl"):
        run = subprocess.run([exe, "--src", "explosion", "--sys", sys, "--stations", str(tmp_path / "st.txt"),
                              "--seis", str(tmp_path / "s.f32"), "--out", str(tmp_path / "o.f32")], capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        res[sys] = np.fromfile(tmp_path / "o.f32", dtype=np.float32).reshape(nrec, 3, ns)
    us, up, uz = (raw[:, :, c].T.astype(float) for c in range(3))          # explosion: all factors 1
    ur = us * np.sin(colat)[:, None] + uz * np.cos(colat)[:, None]
    ut = us * np.cos(colat)[:, None] - uz * np.sin(colat)[:, None]
    tol = 3e-6 * np.abs(raw).max()
    assert np.abs(res["sph"] - np.stack([ut, up, ur], axis=1)).max() < tol
    assert np.abs(res["enz"] - np.stack([-ut, up, ur], axis=1)).max() < tol
    assert np.abs(res["cyl"] - np.stack([us, up, uz], axis=1)).max() < tol


def test_causal_stf_convolution(exe, tmp_path):
    """convolve_with_stf: result = pi * sum_j seis(i-j) stf(j dt) dt with the Gaussian centred at
    1.5 t_0 (so traces come out delayed by that) — an impulse returns the kernel itself."""
    ns, dt, t0 = 400, 0.5, 20.0
    raw = np.zeros((ns, 1, 3), np.float32)
    raw[10, 0, :] = (1.0, 2.0, -1.0)
    raw.tofile(tmp_path / "s.f32")
    (tmp_path / "st.txt").write_text("60.0 0.0\n")
    run = subprocess.run([exe, "--src", "explosion", "--sys", "cyl", "--stf-conv", repr(t0), repr(dt), "gauss_0",
                          "--stations", str(tmp_path / "st.txt"), "--seis", str(tmp_path / "s.f32"),
                          "--out", str(tmp_path / "o.f32")], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    got = np.fromfile(tmp_path / "o.f32", dtype=np.float32).reshape(1, 3, ns)
    a = 3.5 / t0
    nj = int(2 * 1.5 * t0 / dt)
    j = np.arange(1, nj + 1)
    ker = a * np.exp(-(a * (j * dt - 1.5 * t0)) ** 2) / np.sqrt(np.pi) / np.pi * dt * np.pi
    want = np.zeros(ns)
    want[10 + j] = ker                                         # (0-based sample 10 = 1-based 11; i = 11 + j)
    assert np.abs(got[0, 0] - want).max() < 1e-6 * ker.max()
    assert np.abs(got[0, 1] - 2 * want).max() < 2e-6 * ker.max()
    assert abs(ker.sum() - 1.0) < 2e-3                         # unit area: amplitudes are preserved


def _write_simdir(d, src1, src2, mag, raw, names, colat, lon, dt, shift, sc, sl):
    """A run directory as the reference's solver leaves it with USE_NETCDF false (simulation.info in
    the formats of parameters.F90:1410-1465, Data/receiver_names.dat, receiver_pts.dat, *_disp.dat)."""
    os.makedirs(d / "Data")
    ns, nrec = raw.shape[0], raw.shape[1]
    A = lambda s: f"{s:>45s}"[:45]
    f21 = lambda v, s: f"{v:22.7f}{A(s)}\n"
    f22 = lambda v, s: f"{v:20d}{A(s)}\n"
    f23 = lambda v, s: f"{v:>20s}{A(s)}\n"
    txt = (f23("prem_iso", "background model") + f21(dt, "time step [s]") + f22(ns - 1, "number of time steps")
           + f23(src1, "source type") + f23(src2, "source type") + f23("dirac_0", "source time function") + f23("moment", "simtype")
           + f21(20.0, "dominant source period") + f21(24.4, "source depth [km]") + f21(sc, "Source colatitude")
           + f21(sl, "Source longitude") + f"{mag:15.5E}{A('scalar source magnitude')}\n" + f22(nrec, "number of receivers")
           + f22(ns, "length of seismogram [time samples]") + f21(dt, "seismogram sampling [s]") + f22(0, "number of strain dumps")
           + f21(0.0, "strain dump sampling rate [s]") + f22(0, "number of snapshot dumps") + f21(0.0, "snapshot dump sampling rate [s]")
           + f23("cyl", "receiver components ") + f22(0, "  ibeg: beginning gll index for wavefield dumps")
           + f22(4, "iend: end gll index for wavefield dumps") + f21(shift, "source shift factor [s]")
           + f22(round(shift / dt), "source shift factor for deltat") + f22(round(shift / dt), "source shift factor for seis_dt")
           + f22(round(shift / dt), "source shift factor for deltat_coarse") + f23("stations", "receiver file type")
           + f21(0.0, "receiver spacing (0 if not even)") + f"{'F':>20s}{A('use netcdf for wavefield output?')}\n"
           + f22(100, "nelem") + f22(20, "nel_fluid") + f22(2, "nproc"))
    (d / "simulation.info").write_text(txt)
    (d / "Data" / "receiver_names.dat").write_text("".join(f" {n} {10.0 + k} {20.0 + k}\n" for k, n in enumerate(names)))
    (d / "Data" / "receiver_pts.dat").write_text("".join(f" {float(np.rad2deg(c))!r} {float(np.rad2deg(l))!r} {k % 2}\n"
                                                         for k, (c, l) in enumerate(zip(colat, lon))))
    for k, n in enumerate(names):
        cols = raw[:, k, :][:, [0, 2]] if src1 == "monopole" else raw[:, k, :]
        np.savetxt(d / "Data" / f"{n}_disp.dat", cols, fmt="%16.8E")


@pytest.mark.parametrize("sys", ["enz", "sph"])
def test_run_directories_in_the_references_layout(exe, tmp_path, sys):
    """--simdir: the four run directories of a moment-tensor source as the reference's solver (or
    axisem_b200_solver --rundir) writes them give what the raw arrays give; --ascii-out writes the
    processed traces the way post_processing.F90:410-425 does."""
    rng = np.random.default_rng(11)
    nrec, ns, dt, shift = 5, 48, 0.5, 4.0
    sc, sl = np.deg2rad(37.5), np.deg2rad(143.0)
    colat = np.deg2rad(rng.uniform(5, 175, nrec))
    lon = np.deg2rad(rng.uniform(0, 360, nrec))
    names = [f"ST{k:02d}_XX" for k in range(nrec)]
    M_dyncm = np.array([1.2e26, -0.7e26, -0.5e26, 2.1e26, -1.4e26, 0.9e26])
    cmt = tmp_path / "CMTSOLUTION"
    cmt.write_text(" PDE 2011  3 11  5 46 23.00  38.3200  142.3700  24.4 7.2 9.0 TEST EVENT\n"
                   "event name:     TEST\ntime shift:      0.0000\nhalf duration:   0.0000\n"
                   "latitude:       52.5\nlongitude:     143.0\ndepth:          24.4\n"
                   + "".join(f"{n}:      {v:.6e}\n" for n, v in zip(("Mrr", "Mtt", "Mpp", "Mrt", "Mrp", "Mtp"), M_dyncm)))
    runs, args = [], []
    for t, s1, mag, sub in (("mrr", "monopole", 1e20, "MZZ"), ("mtt_p_mpp", "monopole", 2e20, "MXX_P_MYY"),
                            ("mtr", "dipole", 1e20, "MXZ_MYZ"), ("mtp", "quadpole", 0.5e20, "MXY_MXX_M_MYY")):
        raw = rng.standard_normal((ns, nrec, 3)).astype(np.float32)
        if s1 == "monopole":
            raw[:, :, 1] = 0.0
        _write_simdir(tmp_path / sub, s1, t, mag, raw, names, colat, lon, dt, shift, sc, sl)
        runs.append((t, mag, raw))
        args += ["--simdir", str(tmp_path / sub)]
    out = tmp_path / "out.f32"
    run = subprocess.run([exe, "--cmt", str(cmt), "--sys", sys, "--out", str(out), "--ascii-out", str(tmp_path / "POST")] + args,
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    got = np.fromfile(out, dtype=np.float32).reshape(nrec, 3, ns)
    want = _expected(runs, M_dyncm / 1e7, colat, lon, sc, sl, sys)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 3e-6 * scale            # the source location came from simulation.info
    comps = {"enz": "NEZ", "sph": "tpr"}[sys]
    for k, n in enumerate(names):
        for c in range(3):
            tab = np.loadtxt(tmp_path / "POST" / "SEISMOGRAMS" / f"{n}_disp_post_mij_conv0000_{comps[c]}.dat")
            assert tab.shape == (ns, 2)
            assert np.allclose(tab[:, 0], np.arange(ns) * dt - shift) and np.abs(tab[:, 1] - got[k, c]).max() <= 1e-6 * scale
    # a run with NetCDF output is refused, not half-read
    info = (tmp_path / "MZZ" / "simulation.info").read_text().replace("                   F", "                   T")
    (tmp_path / "MZZ" / "simulation.info").write_text(info)
    run = subprocess.run([exe, "--simdir", str(tmp_path / "MZZ"), "--out", str(out)], capture_output=True, text=True)
    assert run.returncode != 0 and "NetCDF" in run.stderr


def test_solver_run_directory_through_the_post_processing(exe, tmp_path):
    """End to end on the files alone: axisem_b200_solver --rundir (its CPU twin) leaves the reference's run
    directory for a source off the pole and STATIONS in geographic coordinates; axisem_b200_postproc --simdir
    reads it — source position from simulation.info, receiver names and solver-frame coordinates from Data/ —
    and returns what it returns for the raw seismogram file of the same run with the rotated station list."""
    from axisem_b200.host import SourceParams, build_problem, prem_mesh_spec
    from axisem_b200.host.meshdb_io import write_meshdb
    from oracle import oracle
    spec = prem_mesh_spec(ntheta=16, nr_target=18)
    prob = build_problem(spec, SourceParams(src_type2="mtr", t_0=3.0), niter=40, seis_it=2)
    db = str(tmp_path / "meshdb.dat0000")
    write_meshdb(prob.mesh, db, dt=prob.deltat)
    st = tmp_path / "STATIONS"
    st.write_text("AAA XX 10.0 20.0 0.0 0.0\nBBB XX -35.5 140.0 0.0 0.0\nCCC YY 62.0 -110.0 0.0 0.0\n")
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp_path / "run"), "--rundir", str(tmp_path / "RUN"), "--src", "mtr",
                          "--period", "3", "--niter", "40", "--seis-it", "2", "--stations", str(st), "--src-lat", "36.5", "--src-lon", "140.25", db],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    a = subprocess.run([exe, "--simdir", str(tmp_path / "RUN"), "--sys", "enz", "--out", str(tmp_path / "a.f32"), "--ascii-out", str(tmp_path / "POST")],
                       capture_output=True, text=True)
    assert a.returncode == 0, a.stderr
    # the reference rotates with the coordinates of the grid points taken (receiver_pts.dat), not those asked for
    pts = np.loadtxt(tmp_path / "RUN" / "Data" / "receiver_pts.dat")
    (tmp_path / "pts.txt").write_text("".join(f"{c!r} {l!r}\n" for c, l in pts[:, :2].tolist()))
    b = subprocess.run([exe, "--src", "mtr", "--seis", str(tmp_path / "run.rank0000.seis.f32"), "--stations", str(tmp_path / "pts.txt"),
                        "--srccolat", repr(90.0 - 36.5), "--srclon", "140.25", "--sys", "enz", "--out", str(tmp_path / "b.f32")],
                       capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    x, y = (np.fromfile(tmp_path / f, dtype=np.float32).reshape(3, 3, -1) for f in ("a.f32", "b.f32"))
    assert x.shape == y.shape == (3, 3, 21) and np.abs(y).max() > 0
    # (simulation.info holds the source position to seven decimals of a radian, the *_disp.dat files nine digits)
    assert np.abs(x - y).max() <= 1e-5 * np.abs(y).max()
    assert sorted(os.listdir(tmp_path / "POST" / "SEISMOGRAMS"))[:3] == ["AAA_XX_disp_post_mij_conv0000_E.dat",
                                                                         "AAA_XX_disp_post_mij_conv0000_N.dat",
                                                                         "AAA_XX_disp_post_mij_conv0000_Z.dat"]
