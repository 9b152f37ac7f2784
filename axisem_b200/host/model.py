"""Background models for the synthetic PREM-type meshes (host side, float64).

Restates the radial polynomials of PREM (Dziewonski & Anderson 1981) exactly as the
reference evaluates them (SOLVER/background_models.F90:417-529 `prem_sub`, :534-674
`prem_ani_sub`) and the conversion to (rho, lambda, mu, xi, phi, eta)
(SOLVER/get_model.F90:160-186).  The ocean/upper-crust layers are merged into one
24.4 km crustal layer (documented deviation: keeps the synthetic meshes' time step sane).

Each layer is (r_bottom_km, r_top_km, fluid?, Qmu, Qkappa, polynomials in x=r/6371).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List

import numpy as np

R_EARTH = 6371.0e3


@dataclass
class Layer:
    r_bot: float        # metres
    r_top: float
    fluid: bool
    qmu: float
    qkappa: float
    rho: Callable[[np.ndarray], np.ndarray]
    vpv: Callable[[np.ndarray], np.ndarray]
    vsv: Callable[[np.ndarray], np.ndarray]
    vph: Callable[[np.ndarray], np.ndarray]
    vsh: Callable[[np.ndarray], np.ndarray]
    eta: Callable[[np.ndarray], np.ndarray]
    name: str = ""


def _poly(*c):
    c = [float(v) for v in c]

    def f(x):
        x = np.asarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        for k, ck in enumerate(c):
            out = out + ck * x ** k
        return out
    return f


_one = _poly(1.0)


def prem_layers(anisotropic: bool = False, r_min_km: float = 400.0) -> List[Layer]:
    """PREM layers from the centre outwards (units SI).  `anisotropic` selects the
    transversely isotropic upper mantle of `prem_ani_sub`; otherwise `prem_sub`'s
    isotropic one.  The sphere is hollow below r_min (free inner surface)."""
    km = 1.0e3
    lay: List[Layer] = []

    def iso(r0, r1, fluid, qmu, qka, rho, vp, vs, name):
        lay.append(Layer(r0 * km, r1 * km, fluid, qmu, qka, rho, vp, vs, vp, vs, _one, name))

    iso(r_min_km, 1221.5, False, 84.6, 1327.7,
        _poly(13.0885, 0, -8.8381), _poly(11.2622, 0, -6.3640), _poly(3.6678, 0, -4.4475),
        "inner core")
    iso(1221.5, 3480.0, True, 0.0, 57827.0,
        _poly(12.5815, -1.2638, -3.6426, -5.5281), _poly(11.0487, -4.0362, 4.8023, -13.5732),
        _poly(0.0), "outer core")
    rho_lm = _poly(7.9565, -6.4761, 5.5283, -3.0807)
    iso(3480.0, 3630.0, False, 312.0, 57827.0, rho_lm,
        _poly(15.3891, -5.3181, 5.5242, -2.5514), _poly(6.9254, 1.4672, -2.0834, 0.9783), "D''")
    iso(3630.0, 5600.0, False, 312.0, 57827.0, rho_lm,
        _poly(24.9520, -40.4673, 51.4832, -26.6419), _poly(11.1671, -13.7818, 17.4575, -9.2777),
        "lower mantle")
    iso(5600.0, 5701.0, False, 312.0, 57827.0, rho_lm,
        _poly(29.2766, -23.6027, 5.5242, -2.5514), _poly(22.3459, -17.2473, -2.0834, 0.9783),
        "lower mantle top")
    iso(5701.0, 5771.0, False, 143.0, 57827.0,
        _poly(5.3197, -1.4836), _poly(19.0957, -9.8672), _poly(9.9839, -4.9324), "TZ 670-600")
    iso(5771.0, 5971.0, False, 143.0, 57827.0,
        _poly(11.2494, -8.0298), _poly(39.7027, -32.6166), _poly(22.3512, -18.5856), "TZ 600-400")
    iso(5971.0, 6151.0, False, 143.0, 57827.0,
        _poly(7.1089, -3.8045), _poly(20.3926, -12.2569), _poly(8.9496, -4.4597), "400-220")
    rho_um = _poly(2.6910, 0.6924)
    for (r0, r1, qmu, name) in ((6151.0, 6291.0, 80.0, "LVZ"), (6291.0, 6346.6, 600.0, "LID")):
        if anisotropic:
            lay.append(Layer(r0 * km, r1 * km, False, qmu, 57827.0, rho_um,
                             _poly(0.8317, 7.2180), _poly(5.8582, -1.4678),
                             _poly(3.5908, 4.6172), _poly(-1.0839, 5.7176),
                             _poly(3.3687, -2.4778), name))
        else:
            iso(r0, r1, False, qmu, 57827.0, rho_um,
                _poly(4.1875, 3.9382), _poly(2.1519, 2.3481), name)
    iso(6346.6, 6356.0, False, 600.0, 57827.0, _poly(2.9), _poly(6.8), _poly(3.9), "lower crust")
    iso(6356.0, 6371.0, False, 600.0, 57827.0, _poly(2.6), _poly(5.8), _poly(3.2), "upper crust")
    return lay


def homogeneous_layers(r_min_km: float = 400.0, r_max_km: float = 6371.0,
                       rho=3000.0, vp=8000.0, vs=4500.0, qmu=300.0, qkappa=57827.0) -> List[Layer]:
    """Single solid layer; for analytic checks."""
    return [Layer(r_min_km * 1e3, r_max_km * 1e3, False, qmu, qkappa,
                  _poly(rho / 1e3), _poly(vp / 1e3), _poly(vs / 1e3), _poly(vp / 1e3),
                  _poly(vs / 1e3), _one, "homogeneous")]


def evaluate_layer(layer: Layer, r: np.ndarray):
    """rho, lambda, mu, xi, phi, eta at radii r (m) — get_model.F90:166-186."""
    x = np.asarray(r, dtype=np.float64) / R_EARTH
    rho = layer.rho(x) * 1.0e3
    vph = layer.vph(x) * 1.0e3
    vpv = layer.vpv(x) * 1.0e3
    vsh = layer.vsh(x) * 1.0e3
    vsv = layer.vsv(x) * 1.0e3
    eta = layer.eta(x)
    lam = rho * (vph ** 2 - 2.0 * vsh ** 2)
    mu = rho * vsh ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        xi = np.where(vsv > 1e-10 * vph, vsh ** 2 / np.where(vsv > 0, vsv, 1.0) ** 2, 1.0)
    phi = vpv ** 2 / vph ** 2
    return rho, lam, mu, xi, phi, eta, np.maximum(vph, vpv)
