"""A MESHER database that is not a theta x r grid: a conforming lateral coarsening ("doubling") layer of
semino / semiso elements between a fine upper and a coarse lower shell (tests/doubling_mesh.py), through the
whole native chain — database reader, element mappings, pre-computation, time loop.

There is no second implementation to compare such a mesh with, so the checks are the reference's own
invariants and a twin run: mass = volume (def_grid.f90:1188), a stable run whose total energy is constant
once the source has acted (the diagnostic of time_evol_wave.F90:1424-1526 — symmetric positive operators on
every element type), and seismograms equal, to discretisation accuracy, to those of the uncoarsened mesh
with the same radial layering."""
import os
import subprocess

import numpy as np
import pytest

from axisem_b200.host.spectral import SpectralBasis

from . import doubling_mesh as dm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRECOMP = os.path.join(ROOT, "axisem_b200", "axisem_b200_precomp")
PRODUCT_EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_solver")
COLAT = "40,80,120,160"
RUN = ["--src", "explosion", "--depth", "300", "--period", "250", "--seis-it", "4", "--receivers", COLAT]


def _databases(tmp_path, nth=32):
    b = SpectralBasis(4)
    out = {}
    for name, flag in (("dbl", True), ("reg", False)):
        M = dm.build(nth=nth, doubling=flag)
        path = str(tmp_path / f"{name}.dat0000")
        dm.write_database(path, M, b, dt=0.5)
        out[name] = (path, M)
    return out


def _traces(rundir, n=4):
    return np.array([np.loadtxt(os.path.join(rundir, "Data", f"recfile_{k:04d}_disp.dat")) for k in range(1, n + 1)])


def test_mesh_is_conforming():
    M = dm.build(nth=32)
    assert M["nelem"] == 8 * 32 + 48 + 3 * 16 and M["ndoubling"] == 48
    assert sorted(set(M["eltype"])) == ["curved", "semino", "semiso"]
    ig = M["igloc"].reshape(M["nelem"], 5, 5)
    # every interior edge is shared by exactly two elements, point for point; boundary edges lie on r_min,
    # router or the axis
    edges = {}
    for e in range(M["nelem"]):
        for pts in (ig[e, 0, :], ig[e, 4, :], ig[e, :, 0], ig[e, :, 4]):
            k = tuple(sorted((int(pts[0]), int(pts[4]))))
            edges.setdefault(k, []).append((e, tuple(int(p) for p in pts)))
    nbound = 0
    for k, users in edges.items():
        assert len(users) in (1, 2)
        if len(users) == 2:
            a, b = users[0][1], users[1][1]
            assert a == b or a == b[::-1]
        else:
            nbound += 1
    assert nbound == 32 + 16 + 2 * 13           # surface, inner surface, the two halves of the axis
    # valence of the template's nodes: P and Q belong to three elements, C to four, and six meet where the
    # diagonals of two periods reach the coarse row
    val = np.bincount(np.concatenate([ig[:, 0, 0], ig[:, 0, 4], ig[:, 4, 0], ig[:, 4, 4]]))
    assert set(val[val > 0]) == {1, 2, 3, 4, 6} and (val == 3).sum() == 16 and (val == 6).sum() == 8


def test_native_chain_on_a_coarsening_layer(tmp_path):
    from oracle import oracle
    db = _databases(tmp_path)
    for name, (path, M) in db.items():
        out = subprocess.run([PRECOMP, "--out", str(tmp_path / f"pre_{name}"), "--niter", "10"] + RUN[:6] + [path],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        checks = dict(line.split() for line in out.stdout.strip().splitlines())
        assert abs(float(checks["mass_over_volume"]) - 1.0) < 1e-9, (name, checks)
        assert int(checks["n_sf_boundaries"]) == 0
    exe = oracle.build_host()
    tr, en = {}, {}
    for name, (path, M) in db.items():
        r = subprocess.run([exe, "--quiet", "--out", str(tmp_path / f"run_{name}"), "--rundir", str(tmp_path / f"RUN_{name}"),
                            "--niter", "4000", "--energy"] + RUN + [path], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr
        tr[name] = _traces(tmp_path / f"RUN_{name}")
        en[name] = np.loadtxt(tmp_path / f"RUN_{name}" / "Data" / "energy_glob.dat")
    for name in db:
        e = en[name]
        late = e[e[:, 0] > 900.0, 3]                    # the source (250 s, centred on 375 s) has acted
        assert late.min() > 0 and (late.max() - late.min()) / late.mean() < 1e-4, name
    assert abs(en["dbl"][-1, 3] / en["reg"][-1, 3] - 1.0) < 1e-3          # the same energy went in
    a, b = tr["dbl"], tr["reg"]
    assert a.shape == b.shape == (4, 1001, 2) and np.abs(b).max() > 0
    # attenuation needs an anelastic model (get_mesh.f90:142-150); prem_iso_light is one, its solid twin is not
    out = subprocess.run([PRECOMP, "--out", str(tmp_path / "x"), "--model", "prem_iso_solid_light", "--attenuation", "cg4", "--niter", "10"]
                         + RUN[:6] + [db["dbl"][0]], capture_output=True, text=True)
    assert out.returncode != 0 and "elastic only" in out.stderr
    for k in range(4):
        assert np.sqrt(((a[k] - b[k]) ** 2).sum() / (b[k] ** 2).sum()) < 0.03, k
        assert np.corrcoef(a[k][:, 1], b[k][:, 1])[0, 1] > 0.999


# ---- the whole sphere: inner square of `linear` elements, a ring of semino / semiso elements around it ----
RC = (1221.5e3, 2350e3, 3480e3, 3630e3, 4115e3, 4600e3)
RF = (4900e3, 5250e3, 5600e3, 5701e3, 5771e3, 5971e3, 6151e3, 6371e3)
DISC = (6371e3, 6151e3, 5971e3, 5771e3, 5701e3, 5600e3, 3630e3, 3480e3, 1221.5e3)


def _sphere_database(tmp_path, name, cube):
    M = dm.build(nth=32, r_coarse=RC, r_fine=RF, cube_halfwidth=cube)
    path = str(tmp_path / f"{name}.dat0000")
    dm.write_database(path, M, SpectralBasis(4), bkgrdmodel="prem_iso_solid_light", discont=DISC, dt=0.5)
    return path, M


def test_native_chain_on_a_full_sphere_with_inner_cube(tmp_path):
    """No hollow centre: the mesher's central square of `linear` (8-node serendipity) elements and the ring
    that joins it to the spherical shell, below a coarsening layer, in prem_iso_solid_light (all solid).
    The volume of the whole sphere, a stable run with constant energy while the wavefield crosses the centre
    several times, and — until anything has reached the inner core — the seismograms of the same mesh
    with a free surface at the inner-core boundary instead."""
    from oracle import oracle
    runs = {"cube": _sphere_database(tmp_path, "cube", 500e3), "hollow": _sphere_database(tmp_path, "hollow", None)}
    M = runs["cube"][1]
    assert {t: M["eltype"].count(t) for t in set(M["eltype"])} == {"linear": 32, "semino": 32, "semiso": 32, "curved": 304}
    assert M["ax_el"].size == 2 * (5 + 1 + 7) + 2 * 4 + 4          # shell + coarsening layer, ring, square
    exe = oracle.build_host()
    tr, en = {}, {}
    for name, (path, _) in runs.items():
        out = subprocess.run([PRECOMP, "--out", str(tmp_path / f"pre_{name}"), "--niter", "10"] + RUN[:6] + [path], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        checks = dict(line.split() for line in out.stdout.strip().splitlines())
        assert abs(float(checks["mass_over_volume"]) - 1.0) < 1e-9, (name, checks)      # cube: the volume of the whole sphere
        r = subprocess.run([exe, "--quiet", "--out", str(tmp_path / f"run_{name}"), "--rundir", str(tmp_path / f"RUN_{name}"),
                            "--niter", "6000", "--energy", "--src", "explosion", "--depth", "300", "--period", "250", "--seis-it", "4",
                            "--receivers", "40,80,120,178", path], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr
        tr[name] = _traces(tmp_path / f"RUN_{name}")
        en[name] = np.loadtxt(tmp_path / f"RUN_{name}" / "Data" / "energy_glob.dat")
        late = en[name][en[name][:, 0] > 900.0, 3]
        assert late.min() > 0 and (late.max() - late.min()) / late.mean() < 1e-4, name
    assert abs(en["cube"][-1, 3] / en["hollow"][-1, 3] - 1.0) < 1e-3
    a, b = tr["cube"], tr["hollow"]
    t = np.arange(a.shape[1]) * 4 * 0.5
    rel = lambda k, w: np.sqrt(((a[k][w] - b[k][w]) ** 2).sum() / (b[k][w] ** 2).sum())
    assert rel(0, t < 700.0) < 1e-4 and rel(1, t < 1200.0) < 1e-3 and rel(2, t < 1200.0) < 1e-3      # mantle phases only
    assert rel(3, t < 3000.0) > 0.05                                                                  # the centre is there


@pytest.mark.gpu
def test_cuda_library_on_a_full_sphere_with_inner_cube(tmp_path):
    from oracle import oracle
    assert os.path.exists(PRODUCT_EXE), "axisem_b200_solver missing: run __graft_entry__.build()"
    path, _ = _sphere_database(tmp_path, "cube", 500e3)
    got = {}
    for name, exe in (("gpu", PRODUCT_EXE), ("cpu", oracle.build_host())):
        r = subprocess.run([exe, "--quiet", "--out", str(tmp_path / name), "--niter", "3000", "--src", "mtr", "--depth", "300",
                            "--period", "250", "--seis-it", "4", "--receivers", "40,80,120,178", path], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr
        got[name] = np.fromfile(tmp_path / f"{name}.rank0000.seis.f32", dtype=np.float32)
    assert got["gpu"].shape == got["cpu"].shape and np.abs(got["cpu"]).max() > 0
    d = got["gpu"].astype(np.float64) - got["cpu"]
    assert np.sqrt((d ** 2).sum() / (got["cpu"].astype(np.float64) ** 2).sum()) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [["--attenuation", "cg4"], ["--attenuation", "full"], ["--scheme", "symplec4"]],
                         ids=["cg4", "full_memvars", "symplec4"])
def test_cuda_library_on_a_coarsening_layer(tmp_path, variant):
    """The device library on the unstructured database (assembly groups of valence 3, 4 and 6 in the layer,
    semino / semiso coefficient planes): the product host against its CPU twin linked to the oracle — Newmark
    with coarse-grained and with full memory variables, and the 4th-order symplectic loop."""
    from oracle import oracle
    assert os.path.exists(PRODUCT_EXE), "axisem_b200_solver missing: run __graft_entry__.build()"
    path, M = _databases(tmp_path)["dbl"]
    got = {}
    for name, exe in (("gpu", PRODUCT_EXE), ("cpu", oracle.build_host())):
        r = subprocess.run([exe, "--quiet", "--out", str(tmp_path / name), "--niter", "2000"] + variant + RUN + [path],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr
        got[name] = np.fromfile(tmp_path / f"{name}.rank0000.seis.f32", dtype=np.float32)
    assert got["gpu"].shape == got["cpu"].shape and np.abs(got["cpu"]).max() > 0
    d = got["gpu"].astype(np.float64) - got["cpu"]
    assert np.sqrt((d ** 2).sum() / (got["cpu"].astype(np.float64) ** 2).sum()) <= 1e-5


# ---- the mesher's domain decomposition of such a mesh ----------------------------------------------------------
EARTH_ROWS = [(1221.5e3, 2350e3, "R"), (2350e3, 3480e3, "R"), (3480e3, 3630e3, "R"), (3630e3, 4115e3, "R"), (4115e3, 4600e3, "R"),
              (4600e3, 4900e3, "D"), (4900e3, 5250e3, "R"), (5250e3, 5600e3, "R"), (5600e3, 5701e3, "R"), (5701e3, 5771e3, "R"),
              (5771e3, 5971e3, "R"), (5971e3, 6151e3, "R"), (6151e3, 6291e3, "R"), (6291e3, 6371e3, "R")]
EARTH_DISC = (6371e3, 6291e3, 6151e3, 5971e3, 5771e3, 5701e3, 5600e3, 3630e3, 3480e3, 1221.5e3)
EARTH_ARGS = ["--src", "mtr", "--depth", "300", "--period", "250", "--seis-it", "4", "--attenuation", "cg4", "--receivers", "40,80,120,178"]


def _earth_databases(tmp_path, tag, nth_blocks, r_cuts):
    F = dm.build_rows(EARTH_ROWS, 16, cube_halfwidth=500e3, fluid=lambda r: 1221.5e3 < r < 3480e3)
    parts = dm.partition(F, nth_blocks, r_cuts)
    files = []
    for r, P in enumerate(parts):
        files.append(str(tmp_path / f"{tag}.dat{r:04d}"))
        dm.write_database(files[-1], P, SpectralBasis(4), bkgrdmodel="prem_iso_light", discont=EARTH_DISC,
                          solid_domain=[1, 1, 1, 1, 1, 1, 1, 1, 0, 1], dt=0.5)
    return files, parts


def _seis_by_station(tmp_path, tag, exe, files, niter, extra=()):
    r = subprocess.run([exe, "--quiet", "--out", str(tmp_path / f"run_{tag}"), "--rundir", str(tmp_path / f"RUN_{tag}"),
                        "--niter", str(niter)] + list(extra) + EARTH_ARGS + files, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    return _traces(tmp_path / f"RUN_{tag}")


@pytest.mark.parametrize("nth_blocks,r_cuts", [(2, (5000e3,)), (3, ())], ids=["2x2_blocks", "3_slices"])
def test_decomposed_whole_earth_equals_one_rank(tmp_path, nth_blocks, r_cuts):
    """One database per rank of a whole Earth (inner square, fluid core, coarsening layer) cut into theta x r
    blocks: every rank of the 2 x 2 case has three neighbours and one point belongs to all four; boundary
    pairs stay on one rank; ranks without fluid, ranks without receivers.  The pre-computation assembles mass
    and boundary terms across the cuts, and the dipole + attenuation run equals the undivided one to the
    order of summation."""
    from oracle import oracle
    one, _ = _earth_databases(tmp_path, "one", 1, ())
    files, parts = _earth_databases(tmp_path, "cut", nth_blocks, r_cuts)
    assert len(files) == nth_blocks * (len(r_cuts) + 1)
    if r_cuts:
        assert all(len(P["halo_solid"][0]) == 3 for P in parts)
        assert sorted(len(l) for l in parts[0]["halo_solid"][1])[0] == 1          # the point where the four blocks meet
        assert any(P["nel_fluid"] == 0 for P in parts) and any(P["nel_fluid"] > 0 for P in parts)
    for P in parts:
        for q, l in zip(*P["halo_solid"]):                                         # symmetric lists
            back = dict(zip(*[parts[q]["halo_solid"][0].tolist(), parts[q]["halo_solid"][1]]))
            assert len(back[P["rank"]]) == len(l)
    out = subprocess.run([PRECOMP, "--out", str(tmp_path / "pre"), "--niter", "10"] + EARTH_ARGS[:6] + files, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    checks = dict(line.split() for line in out.stdout.strip().splitlines())
    assert abs(float(checks["mass_over_volume"]) - 1.0) < 1e-9 and abs(float(checks["bdry_sum"]) - 4.0) < 1e-9
    exe = oracle.build_host()
    a = _seis_by_station(tmp_path, "one", exe, one, 2000)
    b = _seis_by_station(tmp_path, "cut", exe, files, 2000)
    assert np.abs(a).max() > 0 and np.sqrt(((a - b) ** 2).sum() / (a ** 2).sum()) < 1e-5


@pytest.mark.gpu
def test_cuda_library_on_a_decomposed_whole_earth(tmp_path):
    """Four theta x r blocks of the whole Earth as four handles on one device (halo words between
    them), against the undivided run of the CPU twin."""
    from oracle import oracle
    assert os.path.exists(PRODUCT_EXE), "axisem_b200_solver missing: run __graft_entry__.build()"
    one, _ = _earth_databases(tmp_path, "one", 1, ())
    files, _ = _earth_databases(tmp_path, "cut", 2, (5000e3,))
    a = _seis_by_station(tmp_path, "cpu", oracle.build_host(), one, 2000)
    b = _seis_by_station(tmp_path, "gpu", PRODUCT_EXE, files, 2000, extra=["--devices", "1"])
    assert np.abs(a).max() > 0 and np.sqrt(((a - b) ** 2).sum() / (a ** 2).sum()) < 1e-5
