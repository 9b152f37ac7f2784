"""Pins against data the reference itself ships.

The reference's tests hold no array-level vectors for the time loop, but
its TEST04 case ships the background model as the reference sampled it
(TESTING/TEST04_anelastic_anisotropic/model.bm -> tests/golden/prem_ani_model_bm.npz, made by
tests/golden/make_model_fixture.py).  The restated prem_ani polynomials and Q values
(background_models.F90:534-674) that feed every pre-computed coefficient of the synthetic
meshes must reproduce that table."""
import os

import numpy as np

from axisem_b200.host.model import R_EARTH, prem_layers

HERE = os.path.dirname(os.path.abspath(__file__))


def test_prem_ani_matches_the_references_tabulated_model():
    z = np.load(os.path.join(HERE, "golden", "prem_ani_model_bm.npz"))
    t = z["table"]
    assert list(z["columns"]) == ["radius", "rho", "vpv", "vsv", "vph", "vsh", "eta", "qka", "qmu"]
    assert t.shape == (160, 9) and t[0, 0] == 6371000.0 and t[-1, 0] == 0.0
    layers = prem_layers(anisotropic=True, r_min_km=0.0)
    # the table runs from the surface down; a discontinuity radius appears twice, upper side first
    tol = np.array([0.006, 0.006, 0.006, 0.006, 0.006, 6e-6, 0.5, 0.005])   # the table's print precision
    for k, row in enumerate(t):
        r = row[0]
        cands = [L for L in layers if L.r_bot - 1e-6 <= r <= L.r_top + 1e-6]
        assert cands, r
        if len(cands) == 2:
            upper_side = k == 0 or t[k - 1, 0] != r
            L = max(cands, key=lambda L: L.r_top) if upper_side else min(cands, key=lambda L: L.r_top)
        else:
            L = cands[0]
        x = r / R_EARTH
        got = np.array([L.rho(x) * 1e3, L.vpv(x) * 1e3, L.vsv(x) * 1e3, L.vph(x) * 1e3, L.vsh(x) * 1e3,
                        float(L.eta(x)), L.qkappa, L.qmu], dtype=np.float64)
        assert np.all(np.abs(got - row[1:]) <= tol), (r, L.name, got, row[1:])
    # fluid outer core: vs = 0 in the table exactly where the layer is flagged fluid
    for L in layers:
        rm = 0.5 * (L.r_bot + L.r_top)
        near = t[np.argmin(np.abs(t[:, 0] - rm))]
        assert (near[3] == 0.0) == L.fluid, L.name


def test_isotropic_test01_table_is_the_same_model_with_isotropic_columns():
    """TEST01's table (columns radius rho vp vs) is prem_ani too; its vp, vs are the vpv, vsv
    columns of the anisotropic table — a property of the fixtures worth knowing when comparing
    with config 1."""
    z = np.load(os.path.join(HERE, "golden", "prem_ani_model_bm.npz"))
    iso = z["table_iso"]
    t = z["table"]
    assert iso.shape == (160, 4)
    assert np.array_equal(iso[:, 0], t[:, 0]) and np.array_equal(iso[:, 1], t[:, 1])
    assert np.array_equal(iso[:, 2], t[:, 2]) and np.array_equal(iso[:, 3], t[:, 3])
