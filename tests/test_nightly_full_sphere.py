"""The reference's golden seismograms (TESTING/nightly/test_0{1,2,3}/ref_data/axisem.mseed, committed as
tests/golden/nightly_ref_seismograms.npz) against a run on a *whole* Earth through the native chain only:
a mesher-format database with the inner square of `linear` elements, the ring that joins it to the
shells, a solid inner core, the fluid outer core with both solid/fluid boundaries, and two coarsening
layers (32 -> 64 columns in the outer core, 64 -> 128 in the lower mantle; tests/doubling_mesh.py) ->
database reader -> element mappings -> get_model (prem_ani) -> pre-computation -> time loop.

tests/test_nightly_reference.py makes the same comparison on the theta x r meshes of the Python builder,
which leave the centre hollow; that is what limited the small core phases of the explosion case on the far
stations there (correlation 0.921 at the worst trace).  With the centre filled the worst trace
and the amplitude range tighten; what remains is the difference of the two meshes (the reference's
mesh is not available: see tests/nightly_compare.py)."""
import os
import subprocess

import numpy as np
import pytest

from axisem_b200.host.spectral import SpectralBasis
from tests.nightly_compare import compare, stations, to_enz

from . import doubling_mesh as dm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_solver")
PRECOMP = os.path.join(ROOT, "axisem_b200", "axisem_b200_precomp")
T_0 = 70.0
DT = 0.14                   # Courant 0.6 on the 9.4 km lower crust
DISC = (6371e3, 6356e3, 6346.6e3, 6291e3, 6151e3, 5971e3, 5771e3, 5701e3, 5600e3, 3630e3, 3480e3, 1221.5e3)
SOLID = [1] * 10 + [0, 1]


def earth_rows():
    km = 1e3
    R = lambda a, b, n: [(a + (b - a) * k / n, a + (b - a) * (k + 1) / n, "R") for k in range(n)]
    rows = R(850 * km, 1221.5 * km, 1)                                          # inner core above the ring
    rows += R(1221.5 * km, 1500 * km, 1) + [(1500 * km, 1800 * km, "D")]        # outer core, columns doubled
    rows += R(1800 * km, 3480 * km, 6)
    rows += R(3480 * km, 3630 * km, 1) + R(3630 * km, 5400 * km, 8) + [(5400 * km, 5600 * km, "D")]     # lower mantle, doubled
    rows += R(5600 * km, 5701 * km, 1) + R(5701 * km, 5771 * km, 1) + R(5771 * km, 5971 * km, 2) + R(5971 * km, 6151 * km, 2)
    # one element between 220 and 80 km: its GLL point at 104.2 km is where the reference's mesh puts the source
    rows += R(6151 * km, 6291 * km, 1) + R(6291 * km, 6346.6 * km, 1) + R(6346.6 * km, 6356 * km, 1) + R(6356 * km, 6371 * km, 1)
    return rows


def _database(tmp_path, ncol, slices=1):
    """-> (database of rank 0, or the list of all ranks' databases when slices > 1; the undivided mesh)"""
    M = dm.build_rows(earth_rows(), ncol, cube_halfwidth=500e3, fluid=lambda r: 1221.5e3 < r < 3480e3)
    files = []
    for r, P in enumerate(dm.partition(M, slices) if slices > 1 else [M]):
        files.append(str(tmp_path / f"earth.dat{r:04d}"))
        dm.write_database(files[-1], P, SpectralBasis(4), bkgrdmodel="prem_ani", discont=DISC, solid_domain=SOLID, dt=DT)
    return (files if slices > 1 else files[0]), M


def _run(exe, path, out, src, scheme="newmark2"):
    names, lat, lon = stations()
    colat = 90.0 - lat
    dt = DT if scheme == "newmark2" else 1.5 * DT        # the symplectic schemes run at 1.5 times the mesher's step
    shift = np.ceil(1.5 * T_0 / dt) * dt
    niter = int((1800.0 + shift) / dt) + 1
    seis_it = max(1, int(0.8 / dt))
    r = subprocess.run([exe, "--quiet", "--out", out, "--src", src, "--depth", "104.2", "--period", str(T_0), "--niter", str(niter),
                        "--seis-it", str(seis_it), "--scheme", scheme, "--rundir", out + "_RUN",
                        "--receivers", ",".join(f"{c:.6f}" for c in colat)] + (path if isinstance(path, list) else [path]),
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr
    # the run directory holds one file per station, whichever rank recorded it (u_s, u_z for a monopole)
    cols = [np.loadtxt(os.path.join(out + "_RUN", "Data", f"recfile_{k + 1:04d}_disp.dat")) for k in range(colat.size)]
    s = np.zeros((cols[0].shape[0], colat.size, 3))
    for k, c in enumerate(cols):
        s[:, k, :] = c if c.shape[1] == 3 else np.stack([c[:, 0], 0 * c[:, 0], c[:, 1]], axis=1)
    t = np.arange(s.shape[0]) * seis_it * dt - shift
    res = compare(src, to_enz(src, s, np.deg2rad(colat), np.deg2rad(lon)), t, T_0)
    big = res[:, :, 3] > 0.002 * res[:, :, 3].max()
    return res[:, :, 0][big], res[:, :, 2][big], s


def test_whole_earth_database_holds_the_references_invariants(tmp_path):
    path, M = _database(tmp_path, 32)
    assert M["nel_fluid"] > 0 and M["ndoubling"] == 6 * 16 + 6 * 32 and M["eltype"].count("linear") == 8 * 16
    out = subprocess.run([PRECOMP, "--out", str(tmp_path / "pre"), "--src", "explosion", "--depth", "104.2", "--period", "70",
                          "--niter", "10", path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    checks = dict(line.split() for line in out.stdout.strip().splitlines())
    assert abs(float(checks["mass_over_volume"]) - 1.0) < 1e-9          # the volume of the whole sphere (def_grid.f90:1188)
    assert int(checks["n_sf_boundaries"]) == 2 and abs(float(checks["bdry_sum"]) - 4.0) < 1e-9      # def_precomp_terms.f90:2743


def test_explosion_on_a_whole_earth_against_the_references_traces(tmp_path):
    """CPU twin of the native host (linked against the oracle), 256 columns at the surface (5 760 elements,
    13 608 steps), cut into four theta-slices with their own databases: a minute and a half on one core, half
    a minute on four."""
    from oracle import oracle
    path, _ = _database(tmp_path, 64, slices=4)
    cc, amp, s = _run(oracle.build_host(), path, str(tmp_path / "run"), "explosion")
    assert cc.size == 40
    # measured: 0.9974 / 0.9995, amplitude 0.988 - 1.025 (median 1.005).  The hollow 224 x 60 mesh of
    # test_nightly_reference.py: 0.921 / 0.9985, 0.84 - 1.07; the reference's own traces against the
    # independent YSPEC solution it ships: 0.997 - 0.9996.
    assert cc.min() > 0.995 and np.median(cc) > 0.999
    assert 0.97 < amp.min() and amp.max() < 1.04 and abs(np.median(amp) - 1.0) < 0.01
    # and against the independent YSPEC solution of this case that the reference ships (no attenuation, no
    # gravity; N and Z): measured 0.920 / 0.9993 (hollow mesh: 0.920 / 0.9984), median amplitude ratio 1.003
    names, lat, lon = stations()
    colat = np.deg2rad(90.0 - lat)
    shift = np.ceil(1.5 * T_0 / DT) * DT
    t = np.arange(s.shape[0]) * max(1, int(0.8 / DT)) * DT - shift
    res = compare("explosion", to_enz("explosion", s, colat, np.deg2rad(lon)), t, T_0, against="yspec")
    big = res[:, :, 3] > 0.002 * res[:, :, 3].max()
    assert np.median(res[:, :, 0][big]) > 0.999 and res[:, :, 0][big].min() > 0.9 and abs(np.median(res[:, :, 2][big]) - 1.0) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("src", ["explosion", "mtr", "mtp"])
def test_cuda_whole_earth_against_the_references_traces(tmp_path, src):
    """The product host on the GPU, 256 columns at the surface, all three source orders."""
    assert os.path.exists(PRODUCT_EXE), "axisem_b200_solver missing: run __graft_entry__.build()"
    path, _ = _database(tmp_path, 64)
    cc, amp, _ = _run(PRODUCT_EXE, path, str(tmp_path / "run"), src)
    # measured on the CPU twin (the device agrees with it to 1e-5): explosion 0.9974 / 0.9995, amplitude 0.988 - 1.025;
    # mtr 0.9860 / 0.9991, 0.991 - 1.024; mtp 0.9984 / 0.9998, 0.995 - 1.022
    floor, med = {"explosion": (0.995, 0.999), "mtr": (0.98, 0.998), "mtp": (0.995, 0.999)}[src]
    assert cc.min() > floor and np.median(cc) > med, (cc.min(), np.median(cc))
    assert 0.97 < amp.min() and amp.max() < 1.04 and abs(np.median(amp) - 1.0) < 0.012, (amp.min(), amp.max(), np.median(amp))


@pytest.mark.gpu
def test_cuda_symplectic_whole_earth_against_the_references_traces(tmp_path):
    """The 4th-order symplectic loop with its point-wise source time function (compute_stf_t), 1.5 times the
    Newmark step, on the same Earth: the dipole traces come out as with Newmark (on the 128-column mesh the
    two loops score 0.8467 / 0.9948 and 0.8468 / 0.9948 on the CPU twin)."""
    assert os.path.exists(PRODUCT_EXE), "axisem_b200_solver missing: run __graft_entry__.build()"
    path, _ = _database(tmp_path, 64)
    cc, amp, _ = _run(PRODUCT_EXE, path, str(tmp_path / "run"), "mtr", scheme="symplec4")
    assert cc.min() > 0.98 and np.median(cc) > 0.998, (cc.min(), np.median(cc))
    assert 0.97 < amp.min() and amp.max() < 1.04 and abs(np.median(amp) - 1.0) < 0.012, (amp.min(), amp.max(), np.median(amp))
