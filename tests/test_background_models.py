"""The 1-D background models of the native host (axisem_b200/hostcxx/background_models.cpp): every
internal bkgrdmodel of the reference (SOLVER/background_models.F90:69-113) and the `external` model
read from a tabulated .bm file (read_ext_model :2082-2421, get_ext_disc :2618-2784, arbitr_sub_solar
:1983-2078, MESHER/interpolation.f90).

Checks: the published values of IASP91 / AK135 / AK135-F at their discontinuities (Kennett &
Engdahl 1991; Kennett et al. 1995; Montagner & Kennett 1996 — values anyone can look up, not taken
from the reference), the domain radii of MESHER/model_discontinuities.f90, the PREM variants against
PREM itself, and the external reader against the reference's own tabulation of prem_ani / prem_iso
(tests/golden/prem_ani_model_bm.npz, from TESTING/TEST04 and TEST01 model.bm)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_precomp")
COLS = ("r", "idom", "rho", "vpv", "vsv", "vph", "vsh", "eta", "qmu", "qka")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["bash", os.path.join(ROOT, "axisem_b200", "hostcxx", "build.sh")])
    return EXE


def evaluate(exe, name, radii, ext=None):
    """radii: km, as numbers (lower side of a discontinuity) or strings 'r+' (upper side)."""
    cmd = [exe] + (["--ext-model", ext] if ext else []) + ["--model-eval", name, ",".join(str(r) for r in radii)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    head, disc, rows = None, [], []
    for line in out.stdout.strip().splitlines():
        w = line.split()
        if w[0] == "ndisc":
            head = {"ndisc": int(w[1]), "anelastic": bool(int(w[3])), "anisotropic": bool(int(w[5])), "name": w[7]}
        elif w[0] == "discont":
            disc.append((float(w[2]), bool(int(w[4]))))
        else:
            rows.append(dict(zip(COLS, map(float, w))))
    return head, disc, rows


# discont(1:ndisc) [km] of MESHER/model_discontinuities.f90 and which domain is fluid
DISCONT = {
    "prem_iso": ([6371, 6356, 6346.6, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 11),
    "prem_ani": ([6371, 6356, 6346.6, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 11),
    "prem_iso_onecrust": ([6371, 6346.6, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 10),
    "prem_ani_onecrust": ([6371, 6346.6, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 10),
    "prem_iso_light": ([6371, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 9),
    "prem_ani_light": ([6371, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 9),
    "prem_iso_solid": ([6371, 6356, 6346.6, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], None),
    "prem_iso_solid_light": ([6371, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], None),
    "prem_crust20_ocean": ([6371, 6367.59, 6365.33, 6364.04, 6361.19, 6358.09, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 14),
    "prem_crust20_cont": ([6371, 6370.04, 6369.21, 6356.33, 6343.72, 6332.84, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 14),
    "prem_crust20_global": ([6371, 6368.8, 6367.34, 6361.32, 6355.03, 6349.18, 6291, 6151, 5971, 5771, 5701, 5600, 3630, 3480, 1221.5], 14),
    "iasp91": ([6371, 6351, 6336, 6251, 6161, 5961, 5711, 5611, 3631, 3482, 1217], 10),
    "ak135": ([6371, 6351, 6336, 6161, 5961, 5711, 3631, 3479.5, 1217.5], 8),
    "ak135f": ([6371, 6361, 6353, 6291, 6251, 6161, 5961, 5711, 5611, 3631, 3479.5, 1217.5], 11),
}


@pytest.mark.parametrize("name", sorted(DISCONT))
def test_domains_of_every_internal_model(exe, name):
    radii, fluid_dom = DISCONT[name]
    # every discontinuity from both sides, and the middle of every domain
    mids = [0.5 * (a + b) for a, b in zip(radii, radii[1:] + [0.0])]
    head, disc, rows = evaluate(exe, name, radii + [f"{r}+" for r in radii[1:]] + mids + [0.0])
    assert head["ndisc"] == len(radii)
    assert np.allclose([d[0] for d in disc], radii, atol=1e-9)
    assert [k + 1 for k, d in enumerate(disc) if d[1]] == ([fluid_dom] if fluid_dom else [])
    assert head["anisotropic"] == ("_ani" in name or "crust20" in name)
    assert head["anelastic"] == ("solid" not in name)
    n = len(radii)
    lower, upper, mid = rows[:n], rows[n:2 * n - 1], rows[2 * n - 1:3 * n - 1]
    assert [int(r["idom"]) for r in lower] == list(range(1, n + 1))          # r = discont(k): domain k below it
    assert [int(r["idom"]) for r in upper] == list(range(1, n))              # 'r+': the domain above
    assert [int(r["idom"]) for r in mid] == list(range(1, n + 1))
    for r in rows:
        assert 1000.0 < r["rho"] < 14000.0 and 1400.0 < r["vpv"] < 14000.0 and 0.0 <= r["vsv"] < 7400.0, r
        assert r["vph"] > 0 and 0.8 < r["eta"] < 1.2
        fluid = fluid_dom is not None and int(r["idom"]) == fluid_dom
        assert (r["vsv"] == 0.0 and r["qmu"] == 0.0) == fluid, r
        if "_ani" not in name and "crust20" not in name:
            assert r["vph"] == r["vpv"] and r["vsh"] == r["vsv"] and r["eta"] == 1.0
    # density and v_p never decrease downwards across a discontinuity by more than the known LVZ / 220 km
    # features allow: a coarse guard against a mis-ordered table
    for up, lo in zip(upper, lower[1:]):
        assert lo["rho"] > 0.9 * up["rho"], (name, up, lo)


def test_published_values_iasp91_ak135(exe):
    def at(name, r):
        return evaluate(exe, name, [r])[2][0]

    def close(v, ref, tol=2e-4):
        return abs(v / ref - 1.0) < tol
    # IASP91 (Kennett & Engdahl 1991, table): Moho 35 km, 410, 660, CMB 2889 km, ICB 5153.9 km, centre
    m = at("iasp91", 6336)
    assert close(m["vpv"], 8040.0) and close(m["vsv"], 4470.0)
    a, b = at("iasp91", "5961+"), at("iasp91", 5961)
    assert close(a["vpv"], 9030.0) and close(a["vsv"], 4870.0) and close(b["vpv"], 9360.0) and close(b["vsv"], 5070.0)
    a, b = at("iasp91", "5711+"), at("iasp91", 5711)
    assert close(a["vpv"], 10200.0) and close(a["vsv"], 5600.0) and close(b["vpv"], 10790.0) and close(b["vsv"], 5950.0)
    a, b = at("iasp91", "3482+"), at("iasp91", 3482)
    assert close(a["vpv"], 13690.8) and close(a["vsv"], 7301.5) and close(b["vpv"], 8008.8) and b["vsv"] == 0
    a, b = at("iasp91", "1217+"), at("iasp91", 1217)
    assert close(a["vpv"], 10257.8) and close(b["vpv"], 11091.4) and close(b["vsv"], 3438.5)
    c = at("iasp91", 0.0)
    assert close(c["vpv"], 11240.9) and close(c["vsv"], 3564.5)
    # AK135 (Kennett, Engdahl & Buland 1995): crust, 410, 660, CMB 2891.5 km, ICB 5153.5 km, centre
    s = at("ak135", 6371)
    assert (s["vpv"], s["vsv"], s["rho"]) == (5800.0, 3460.0, 2720.0)
    a, b = at("ak135", "5961+"), at("ak135", 5961)
    assert close(a["vpv"], 9030.0) and close(a["vsv"], 4870.0) and close(b["vpv"], 9360.0) and close(b["vsv"], 5080.0)
    a, b = at("ak135", "5711+"), at("ak135", 5711)
    assert close(a["vpv"], 10200.0) and close(a["vsv"], 5610.0)
    a, b = at("ak135", "3479.5+"), at("ak135", 3479.5)
    assert close(a["vpv"], 13660.1) and close(a["vsv"], 7281.7, 1e-3) and close(b["vpv"], 8000.0, 6e-3)    # cubic fit of the core
    b = at("ak135", 1217.5)
    assert close(b["vpv"], 11042.7, 1e-3) and close(b["vsv"], 3504.3, 1e-3)
    c = at("ak135", 0.0)
    assert close(c["vpv"], 11262.2, 1e-3) and close(c["vsv"], 3667.8, 1e-3)
    # AK135-F (Montagner & Kennett 1996): radius-dependent Q, the 80 / 120 / 210 km structure
    for r, vp, vs in [("6291+", 8040.0, 4480.0), (6291, 8045.0, 4490.0), (6251, 8050.5, 4500.0), (6161, 8300.5, 4518.9),
                      ("5961+", 9030.3, 4869.8), (5961, 9360.1, 5080.5), ("3479.5+", 13660.1, 7281.7), (3479.5, 7990.4, None)]:
        v = at("ak135f", r)
        assert close(v["vpv"], vp) and (vs is None or close(v["vsv"], vs)), (r, v)
    v = at("ak135f", 6291)
    assert close(v["qmu"], 75.6, 1e-3) and close(v["qka"], 182.03, 1e-3)          # the low-Q asthenosphere
    assert close(at("ak135f", 0.0)["qmu"], 85.03) and close(at("ak135f", 3000.0)["qka"], 57822.0)


def test_prem_variants_are_prem(exe):
    radii = [6300.0, 6200.0, 6000.0, 5800.0, 5750.0, 5650.0, 4500.0, 3550.0, 2500.0, 600.0]
    keys = ("rho", "vpv", "vsv", "vph", "vsh", "eta", "qmu", "qka")
    for base, variants in (("prem_iso", ["prem_iso_onecrust", "prem_iso_light"]),
                           ("prem_ani", ["prem_ani_onecrust", "prem_ani_light", "prem_crust20_ocean", "prem_crust20_global"])):
        ref = evaluate(exe, base, radii)[2]
        for v in variants:
            got = evaluate(exe, v, radii)[2]
            for a, b in zip(ref, got):
                assert all(a[k] == b[k] for k in keys), (v, a, b)
    # one crustal layer = the upper crust down to the Moho; no crust = the LID up to the surface
    one = evaluate(exe, "prem_iso_onecrust", [6350.0])[2][0]
    assert (one["rho"], one["vpv"], one["vsv"]) == (2600.0, 5800.0, 3200.0)
    lid = evaluate(exe, "prem_iso", [6300.0, 6365.0])[2][0]
    light = evaluate(exe, "prem_iso_light", [6365.0])[2][0]
    x = 6365.0 / 6371.0
    assert abs(light["vpv"] - (4.1875 + 3.9382 * x) * 1000) < 1e-9 and light["idom"] == 1 and lid["idom"] == 3
    # the solid variants: a solid outer core with v_s = v_p / sqrt(3), elsewhere PREM
    ref = evaluate(exe, "prem_iso", radii)[2]
    for v in ("prem_iso_solid", "prem_iso_solid_light"):
        got = evaluate(exe, v, radii)[2]
        for a, b in zip(ref, got):
            if a["vsv"] == 0.0:
                assert b["vpv"] == a["vpv"] and abs(b["vsv"] - a["vpv"] / np.sqrt(3.0)) < 1e-9 and b["rho"] == a["rho"]
            else:
                assert all(a[k] == b[k] for k in ("rho", "vpv", "vsv")), (v, a, b)


# ---- external model ------------------------------------------------------------------------------------------
def _fixture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "prem_ani_model_bm.npz"))
    return z["table"], z["table_iso"]


def _write_bm(path, table, columns, units="m", anel=True, ani=True, flip=False, depth=False, name="prem_ani"):
    t = np.array(table, dtype=np.float64)
    if depth:
        t[:, 0] = t[:, 0].max() - t[:, 0]
    if units == "km":
        t[:, 0] /= 1000.0
    if flip:
        t = t[::-1]
    with open(path, "w") as f:
        f.write("# written by tests/test_background_models.py in the format of the reference's model.bm files\n")
        f.write(f"NAME         {name}\nANELASTIC       {'T' if anel else 'F'}\nANISOTROPIC     {'T' if ani else 'F'}\nUNITS        {units}\n")
        f.write("COLUMNS  " + "  ".join(columns) + "\n")
        prev = None
        for row in t:
            if prev is not None and row[0] == prev:
                f.write("#          Discontinuity\n")
            f.write("  ".join(f"{v:.9g}" for v in row) + "\n")
            prev = row[0]


ANI_COLS = ["radius", "rho", "vpv", "vsv", "vph", "vsh", "eta", "qka", "qmu"]


def test_external_model_is_the_reference_tabulation_of_prem(exe, tmp_path):
    table, table_iso = _fixture()
    bm = str(tmp_path / "prem_ani.bm")
    _write_bm(bm, table, ANI_COLS)
    r_nodes = table[:, 0] / 1000.0
    # first-order discontinuities = repeated radii: PREM's own twelve domains
    head, disc, _ = evaluate(exe, "external", [100.0], ext=bm)
    assert head == {"ndisc": 12, "anelastic": True, "anisotropic": True, "name": "prem_ani"}
    assert np.allclose([d[0] for d in disc], DISCONT["prem_ani"][0], atol=1e-9)
    assert [k + 1 for k, d in enumerate(disc) if d[1]] == [11]
    # at the nodes the table itself (through single precision), on the side the row belongs to
    rep = np.isclose(r_nodes[1:], r_nodes[:-1])
    upper_side = np.concatenate([rep, [False]])          # the first row of a repeated pair is the bottom of the domain above
    q = [f"{r:.10g}+" if up else f"{r:.10g}" for r, up in zip(r_nodes, upper_side)]
    rows = evaluate(exe, "external", q[1:], ext=bm)[2]   # (the surface row: r = router is on the first domain's top)
    names = {"rho": 1, "vpv": 2, "vsv": 3, "vph": 4, "vsh": 5, "eta": 6, "qka": 7, "qmu": 8}
    for row, ref in zip(rows, table[1:]):
        for k, c in names.items():
            assert row[k] == float(np.float32(ref[c])), (row, ref, k)
    # between the nodes: linear interpolation of a table of the polynomials = the polynomials to the
    # table's resolution
    rng = np.random.default_rng(5)
    rr = np.sort(rng.uniform(10.0, 6370.0, 300))[::-1]
    rr = rr[np.min(np.abs(rr[:, None] - np.array(DISCONT["prem_ani"][0])[None, :]), axis=1) > 0.5]
    ext = evaluate(exe, "external", list(rr), ext=bm)[2]
    ana = evaluate(exe, "prem_ani", list(rr))[2]
    for a, b in zip(ext, ana):
        assert a["idom"] == b["idom"]
        for k in ("rho", "vpv", "vph", "eta", "qmu", "qka"):
            assert abs(a[k] - b[k]) <= 2e-4 * abs(b[k]), (k, a, b)
        assert abs(a["vsv"] - b["vsv"]) <= 2e-4 * b["vpv"] and abs(a["vsh"] - b["vsh"]) <= 2e-4 * b["vpv"]
    # the same model written in km, in depth and from the centre outwards reads the same
    bm2 = str(tmp_path / "prem_ani_depth_km.bm")
    cols2 = ["depth"] + ANI_COLS[1:]
    _write_bm(bm2, table, cols2, units="km", flip=True, depth=True)
    ext2 = evaluate(exe, "external", list(rr), ext=bm2)[2]
    for a, b in zip(ext, ext2):
        assert a["idom"] == b["idom"]
        for k in names:
            assert abs(a[k] - b[k]) <= 1e-6 * abs(a[k]) + 1e-9, (k, a, b)        # km -> m through single precision
    # the isotropic, elastic table of TEST01: vph = vpv, vsh = vsv, eta = 1; Q is refused as in the reference
    bm3 = str(tmp_path / "prem_iso.bm")
    _write_bm(bm3, table_iso, ["radius", "rho", "vp", "vs"], anel=False, ani=False, name="prem_iso_el")
    head, _, iso = evaluate(exe, "external", list(rr), ext=bm3)
    assert head["ndisc"] == 12 and not head["anelastic"] and not head["anisotropic"]
    for a, b in zip(iso, ana):      # (TEST01's table holds the vertical velocities of prem_ani)
        assert a["vph"] == a["vpv"] and a["vsh"] == a["vsv"] and a["eta"] == 1.0 and a["qmu"] == 0.0
        assert abs(a["vpv"] - b["vpv"]) <= 2e-4 * b["vpv"] and abs(a["rho"] - b["rho"]) <= 2e-4 * b["rho"]
        assert abs(a["vsv"] - b["vsv"]) <= 2e-4 * b["vpv"]


def test_external_model_discontinuity_detection_and_errors(exe, tmp_path):
    # a smooth model with one kink (second-order discontinuity: gradient step >= 0.1 /s) and no jump
    bm = str(tmp_path / "kink.bm")
    r = np.array([1000.0, 800.0, 600.0, 400.0, 200.0, 0.0]) * 1000.0       # metres; written in km
    vp = np.array([3000.0, 3100.0, 3200.0, 3300.0, 3300.0 + 200e3 * 0.2, 3300.0 + 400e3 * 0.2])
    tab = np.stack([r, np.full(6, 2000.0), vp, vp / 2], axis=1)
    _write_bm(bm, tab, ["radius", "rho", "vp", "vs"], units="km", anel=False, ani=False, name="kink")
    head, disc, rows = evaluate(exe, "external", [900.0, 300.0, 100.0], ext=bm)
    assert head["ndisc"] == 2 and [d[0] for d in disc] == [1000.0, 400.0]
    assert [int(x["idom"]) for x in rows] == [1, 2, 2]
    assert abs(rows[0]["vpv"] - 3050.0) < 1e-6 and abs(rows[1]["vpv"] - (3300.0 + 100e3 * 0.2)) < 1e-3
    # no discontinuity at all: a blind one in the middle of the table
    bm = str(tmp_path / "smooth.bm")
    tab[:, 2] = np.linspace(3000.0, 4000.0, 6)
    tab[:, 3] = tab[:, 2] / 2
    _write_bm(bm, tab, ["radius", "rho", "vp", "vs"], units="km", anel=False, ani=False)
    head, disc, _ = evaluate(exe, "external", [10.0], ext=bm)
    assert head["ndisc"] == 2 and disc[1][0] == 600.0
    # what the reference stops on
    for bad, msg in (("NAME x\nANISOTROPIC F\nUNITS m\nCOLUMNS radius rho vp vs\n6371000 1 1 1\n0 1 1 1\n", "ANELASTIC"),
                     ("ANELASTIC F\nANISOTROPIC F\nUNITS m\nCOLUMNS radius rho vp\n6371000 1 1\n0 1 1\n", "vsv"),
                     ("ANELASTIC F\nANISOTROPIC F\nUNITS m\nCOLUMNS radius rho vp vs\n6371 1 1 1\n0 1 1 1\n", "UNITS km"),
                     ("ANELASTIC F\nANISOTROPIC F\nUNITS km\nCOLUMNS radius rho vp vs\n6371 1 1 1\n5000 1 1 1\n5500 1 1 1\n0 1 1 1\n", "monoton")):
        p = str(tmp_path / "bad.bm")
        open(p, "w").write(bad)
        out = subprocess.run([exe, "--ext-model", p, "--model-eval", "external", "100"], capture_output=True, text=True)
        assert out.returncode != 0 and msg in out.stderr, (msg, out.stderr)
    out = subprocess.run([exe, "--model-eval", "external", "100"], capture_output=True, text=True)
    assert out.returncode != 0 and "--ext-model" in out.stderr
