"""Pre-computed per-GLL-point terms ("M-matrices") for the synthetic meshes.

Host-side, vectorised numpy restatement of SOLVER/def_precomp_terms.f90 for concentric
spheroidal elements: it produces the immutable arrays the time loop reads
(SOLVER/data_matr.f90:36-112, data_pointwise.f90), i.e. the *inputs* of the hot path.
Reference sections followed:

  pointwise-derivative matrices   def_precomp_terms.f90:148-304
  mass matrices                   def_precomp_terms.f90:596-835
  solid stiffness terms           def_precomp_terms.f90:1166-1628 + :1632-1812 (monopole),
                                  :1816-2060 (dipole), :2064-2280 (quadrupole)
  TI elastic tensor               def_precomp_terms.f90:2284-2332
  fluid stiffness terms           def_precomp_terms.f90:2336-2470
  solid/fluid boundary terms      def_precomp_terms.f90:2474-2710
  mapping derivatives             analytic_spheroid_mapping.f90:40-107,
                                  analytic_mapping.f90:71-570

All arrays are returned in the reference's memory order: a Fortran `A(0:4,0:4,nel)` is a
C-contiguous numpy array of shape (nel, 5[jpol], 5[ipol]); `A0(0:4,nel)` is (nel, 5).
Everything is evaluated in float64 and rounded once to float32, as the Fortran does on
assignment to its `realkind` arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from .mesh import ElementSet, LocalMesh, MeshSpec, element_coords, make_elements
from .model import evaluate_layer
from .spectral import SpectralBasis

SRC_ORDER = {"monopole": 0, "dipole": 1, "quadpole": 2}


# --------------------------------------------------------------------------------------
@dataclass
class Geometry:
    """Mapping derivatives and quadrature factors at every point of an element set."""
    xi: np.ndarray
    th: np.ndarray
    r: np.ndarray
    s: np.ndarray
    z: np.ndarray
    dsdxi: np.ndarray
    dzdxi: np.ndarray
    dsdeta: np.ndarray
    dzdeta: np.ndarray
    jac: np.ndarray
    W: np.ndarray            # s * w_i w_j  (axial: s/(1+xi) * wax_i w_j)
    W2: np.ndarray           # w_i w_j      (axial: wax_i w_j/(1+xi), 0 at i=0)
    massmat_k: np.ndarray
    massmat_kwts2: np.ndarray
    sin_t: np.ndarray
    cos_t: np.ndarray


def geometry(es: ElementSet, basis: SpectralBasis) -> Geometry:
    xi, th, r, s, z, sin_t, cos_t = element_coords(es, basis)
    nel = es.nel
    dth = 0.5 * (es.th_b - es.th_a)
    dr = 0.5 * (es.r_b - es.r_a)
    R = r[:, :, None]
    ST = sin_t[:, None, :]
    CT = cos_t[:, None, :]
    # analytic_spheroid_mapping.f90:75-104 for concentric elements
    dsdxi = R * CT * dth[:, None, None]
    dzdxi = -R * ST * dth[:, None, None]
    dsdeta = dr[:, None, None] * ST * np.ones_like(R)
    dzdeta = dr[:, None, None] * CT * np.ones_like(R)
    jac = dsdxi * dzdeta - dsdeta * dzdxi
    ax = es.axis
    wxi = np.where(ax[:, None], basis.wt_axial_k[None, :], basis.wt[None, :])   # (nel,5)
    ww = basis.wt[None, :, None] * wxi[:, None, :]                               # (nel,j,i)
    # s/(1+xi) with L'Hospital on the axis (analytic_mapping.f90:71-93)
    opx = 1.0 + xi
    with np.errstate(divide="ignore", invalid="ignore"):
        sop = s / opx[:, None, :]
    sop[:, :, 0] = np.where(ax[:, None], dsdxi[:, :, 0], sop[:, :, 0])
    W = np.where(ax[:, None, None], sop, s) * ww
    with np.errstate(divide="ignore"):
        inv_opx = np.where(opx > 0, 1.0 / np.where(opx > 0, opx, 1.0), 0.0)
    W2 = ww * np.where(ax[:, None, None], inv_opx[:, None, :], 1.0)
    massmat_k = jac * W
    # massmat_kwts2: def_precomp_terms.f90:673-707
    with np.errstate(divide="ignore", invalid="ignore"):
        m2_non = jac / s * ww
        m2_ax = jac / (s * opx[:, None, :]) * ww
    m2_ax[:, :, 0] = (jac[:, :, 0] / np.where(ax[:, None], sop[:, :, 0], 1.0)) * ww[:, :, 0]
    massmat_kwts2 = np.where(ax[:, None, None], m2_ax, m2_non)
    return Geometry(xi, th, r, s, z, dsdxi, dzdxi, dsdeta, dzdeta, jac, W, W2,
                    massmat_k, massmat_kwts2, sin_t, cos_t)


def material(spec: MeshSpec, es: ElementSet, g: Geometry):
    """rho, lambda, mu, xi, phi, eta, vp and Q per element point (get_model.F90:155-186);
    each element takes the polynomial of its own layer, so discontinuities are sharp."""
    shape = g.s.shape
    out = [np.zeros(shape) for _ in range(7)]
    qmu = np.zeros(es.nel)
    qka = np.zeros(es.nel)
    rr = np.broadcast_to(g.r[:, :, None], shape)
    for k, L in enumerate(spec.layers):
        m = es.layer == k
        if not m.any():
            continue
        vals = evaluate_layer(L, rr[m])
        for o, v in zip(out, vals):
            o[m] = v
        qmu[m] = L.qmu
        qka[m] = L.qkappa
    rho, lam, mu, xi_a, phi_a, eta_a, vp = out
    return rho, lam, mu, xi_a, phi_a, eta_a, vp, qmu, qka


def c_ijkl_ani(lam, mu, xi_ani, phi_ani, eta_ani, sin_fa, cos_fa, i, j, k, l, _cache=None):
    """def_precomp_terms.f90:2284-2332 with fast axis s=(sin th, 0, cos th) (radial TI:
    get_model.F90:185-186).  Terms whose Kronecker/fast-axis factor vanishes identically
    are skipped (they contribute exact zeros in the reference)."""
    d = lambda a, b: 1.0 if a == b else 0.0
    i, j, k, l = i - 1, j - 1, k - 1, l - 1
    c = 0.0
    f = d(i, j) * d(k, l)
    if f:
        c = c + lam * f
    f = d(i, k) * d(j, l) + d(i, l) * d(j, k)
    if f:
        c = c + mu * f
    if _cache is not None and _cache.get("iso", False):
        return c + 0.0 * lam if np.isscalar(c) else c
    s = (sin_fa, None, cos_fa)           # s[1] == 0

    def ss(a, b):
        if a == 1 or b == 1:
            return None
        return s[a] * s[b]

    if _cache is None:
        _cache = {}
    if "A" not in _cache:
        _cache["A"] = (eta_ani - 1.0) * lam + 2.0 * eta_ani * mu * (1.0 - 1.0 / xi_ani)
        _cache["B"] = mu * (1.0 / xi_ani - 1.0)
        _cache["C"] = ((1.0 - 2.0 * eta_ani + phi_ani) * (lam + 2.0 * mu)
                       + (4.0 * eta_ani - 4.0) * mu / xi_ani)
    t = None
    for (dd, a, b) in ((d(i, j), k, l), (d(k, l), i, j)):
        if dd:
            v = ss(a, b)
            if v is not None:
                t = v * dd if t is None else t + v * dd
    if t is not None:
        c = c + _cache["A"] * t
    t = None
    for (dd, a, b) in ((d(i, k), j, l), (d(i, l), j, k), (d(j, k), i, l), (d(j, l), i, k)):
        if dd:
            v = ss(a, b)
            if v is not None:
                t = v * dd if t is None else t + v * dd
    if t is not None:
        c = c + _cache["B"] * t
    if 1 not in (i, j, k, l):
        c = c + _cache["C"] * (s[i] * s[j] * s[k] * s[l])
    if np.isscalar(c):
        c = c + 0.0 * lam
    return c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------------------
def solid_stiffness_terms(src_type: str, es: ElementSet, g: Geometry, lam, mu, xi_a, phi_a,
                          eta_a, anel: bool, basis: SpectralBasis) -> Dict[str, np.ndarray]:
    """All `M*` planes of data_matr.f90:46-76 for one source order."""
    ij = 1.0 / g.jac
    W, W2 = g.W, g.W2
    # analytic_mapping.f90:114-570 (the s factor and quadrature weights are inside W)
    alpha = -ij * g.dsdxi * g.dsdeta * W
    beta = ij * g.dsdxi ** 2 * W
    gamma = ij * g.dsdeta ** 2 * W
    delta = -ij * g.dzdxi * g.dzdeta * W
    epsil = ij * g.dzdxi ** 2 * W
    zeta = ij * g.dzdeta ** 2 * W
    Ms_ze_sx = ij * g.dsdxi * g.dzdeta * W
    Ms_ze_se = -ij * g.dsdeta * g.dzdeta * W
    Ms_zx_se = ij * g.dsdeta * g.dzdxi * W
    Ms_zx_sx = -ij * g.dsdxi * g.dzdxi * W
    M_s_xi = g.dsdxi * W2
    M_z_xi = -g.dzdxi * W2
    M_z_eta = g.dzdeta * W2
    M_s_eta = -g.dsdeta * W2
    kw2 = g.massmat_kwts2
    ax = es.axis
    ax3 = ax[:, None, None]

    ST = np.broadcast_to(g.sin_t[:, None, :], lam.shape)
    CT = np.broadcast_to(g.cos_t[:, None, :], lam.shape)
    iso = bool(np.all(xi_a == 1.0) and np.all(phi_a == 1.0) and np.all(eta_a == 1.0))
    cc = {"iso": iso}
    C = lambda a, b, c, d: c_ijkl_ani(lam, mu, xi_a, phi_a, eta_a, ST, CT, a, b, c, d, cc)
    C11, C12, C13, C15 = C(1, 1, 1, 1), C(1, 1, 2, 2), C(1, 1, 3, 3), C(1, 1, 3, 1)
    C22, C23, C25 = C(2, 2, 2, 2), C(2, 2, 3, 3), C(2, 2, 3, 1)
    C33, C35 = C(3, 3, 3, 3), C(3, 3, 3, 1)
    C44, C46, C55, C66 = C(2, 3, 2, 3), C(2, 3, 1, 2), C(3, 1, 3, 1), C(1, 2, 1, 2)

    # axial vectors live at ipol = 0 : shape (nel, 5[jpol])
    ndf = kw2[:, :, 0]                       # non_diag_fact, def_precomp_terms.f90:1316-1329
    w0 = basis.wt_axial_k[0] * basis.wt[None, :]
    dsdxi0 = g.dsdxi[:, :, 0]
    dzdxi0 = g.dzdxi[:, :, 0]
    a0 = lambda A: A[:, :, 0]
    axm = ax[:, None]

    out: Dict[str, np.ndarray] = {}

    def zero_axis(name_list):
        for n in name_list:
            out[n] = np.where(ax3 & (np.arange(5)[None, None, :] == 0), 0.0, out[n])

    if src_type == "monopole":
        out["M11s"] = C11 * delta + C15 * Ms_ze_sx + C15 * Ms_zx_se + C55 * alpha
        out["M21s"] = C11 * zeta + C15 * 2.0 * Ms_ze_se + C55 * gamma
        out["M41s"] = C11 * epsil + C15 * 2.0 * Ms_zx_sx + C55 * beta
        out["M12s"] = C15 * delta + C13 * Ms_ze_sx + C55 * Ms_zx_se + C35 * alpha
        out["M22s"] = C15 * zeta + (C13 + C55) * Ms_ze_se + C35 * gamma
        out["M32s"] = C15 * delta + C13 * Ms_zx_se + C55 * Ms_ze_sx + C35 * alpha
        out["M42s"] = C15 * epsil + (C13 + C55) * Ms_zx_sx + C35 * beta
        out["M11z"] = C55 * delta + C35 * Ms_ze_sx + C35 * Ms_zx_se + C33 * alpha
        out["M21z"] = C55 * zeta + C35 * 2.0 * Ms_ze_se + C33 * gamma
        out["M41z"] = C55 * epsil + C35 * 2.0 * Ms_zx_sx + C33 * beta
        out["M_1"] = C12 * M_z_eta + C25 * M_s_eta
        out["M_2"] = C12 * M_z_xi + C25 * M_s_xi
        out["M_3"] = C23 * M_s_eta + C25 * M_z_eta
        out["M_4"] = C23 * M_s_xi + C25 * M_z_xi
        out["M_w1"] = C22 * kw2
        zero_axis(["M_w1"])
        out["M0_w1"] = np.where(axm, (2.0 * a0(C12) + a0(C22)) * ndf, 0.0)
        out["M0_w2"] = np.where(axm, a0(C25) * ndf, 0.0)
        out["M0_w3"] = np.where(axm, a0(C23) * dsdxi0 * w0 - a0(C25) * dzdxi0 * w0, 0.0)
    elif src_type == "dipole":
        sum_ms = Ms_ze_sx + Ms_zx_se
        out["M11s"] = (C11 + C66) * delta + (C15 + C46) * sum_ms + (C55 + C44) * alpha
        out["M21s"] = (C11 + C66) * zeta + (C15 + C46) * 2.0 * Ms_ze_se + (C55 + C44) * gamma
        out["M41s"] = (C11 + C66) * epsil + (C15 + C46) * 2.0 * Ms_zx_sx + (C55 + C44) * beta
        out["M12s"] = (C11 - C66) * delta + (C15 - C46) * sum_ms + (C55 - C44) * alpha
        out["M22s"] = (C11 - C66) * zeta + (C15 - C46) * 2.0 * Ms_ze_se + (C55 - C44) * gamma
        out["M42s"] = (C11 - C66) * epsil + (C15 - C46) * 2.0 * Ms_zx_sx + (C55 - C44) * beta
        out["M13s"] = C15 * delta + C13 * Ms_ze_sx + C55 * Ms_zx_se + C35 * alpha
        out["M32s"] = C15 * zeta + (C13 + C55) * Ms_ze_se + C35 * gamma
        out["M33s"] = C15 * delta + C13 * Ms_zx_se + C55 * Ms_ze_sx + C35 * alpha
        out["M43s"] = C15 * epsil + (C13 + C55) * Ms_zx_sx + C35 * beta
        out["M11z"] = C55 * delta + C35 * sum_ms + C33 * alpha
        out["M21z"] = C55 * zeta + C35 * 2.0 * Ms_ze_se + C33 * gamma
        out["M41z"] = C55 * epsil + C35 * 2.0 * Ms_zx_sx + C33 * beta
        out["M_1"] = (C12 + C66) * 2.0 * M_z_eta + (C25 + C46) * 2.0 * M_s_eta
        out["M_2"] = (C12 + C66) * 2.0 * M_z_xi + (C25 + C46) * 2.0 * M_s_xi
        out["M_3"] = C46 * M_z_eta + C44 * M_s_eta
        out["M_4"] = C46 * M_z_xi + C44 * M_s_xi
        out["M_5"] = (C12 - C66) * 2.0 * M_z_eta + (C25 - C46) * 2.0 * M_s_eta
        out["M_6"] = (C12 - C66) * 2.0 * M_z_xi + (C25 - C46) * 2.0 * M_s_xi
        out["M_7"] = C25 * 2.0 * M_z_eta + C23 * 2.0 * M_s_eta
        out["M_8"] = C25 * 2.0 * M_z_xi + C23 * 2.0 * M_s_xi
        out["M_w1"] = 4.0 * (C22 + C66) * kw2
        out["M_w2"] = 2.0 * C46 * kw2
        out["M_w3"] = C44 * kw2
        zero_axis(["M_1", "M_2", "M_3", "M_4", "M_5", "M_6", "M_7", "M_8", "M_w1", "M_w3"])
        out["M0_w1"] = np.where(axm, (a0(C12) + a0(C66)) * 2.0 * ndf, 0.0)
        out["M0_w2"] = np.where(axm, -(a0(C12) + a0(C66)) * 2.0 * dzdxi0 * w0, 0.0)
        out["M0_w3"] = np.where(axm, a0(C46) * ndf, 0.0)
        out["M0_w4"] = np.where(axm, -a0(C46) * dzdxi0 * w0, 0.0)
        out["M0_w5"] = np.zeros_like(ndf)
        out["M0_w6"] = np.where(axm, (a0(C25) + a0(C46)) * 2.0 * dsdxi0 * w0, 0.0)
        out["M0_w7"] = np.where(axm, a0(C44) * ndf, 0.0)
        out["M0_w8"] = np.where(axm, a0(C44) * dsdxi0 * w0, 0.0)
        out["M0_w9"] = np.where(axm, (a0(C12) + a0(C22)) * 4.0 * ndf, 0.0)
        out["M0_w10"] = np.where(axm, (2.0 * a0(C25) + a0(C46)) * ndf, 0.0)
    elif src_type == "quadpole":
        out["M11s"] = C11 * delta + C15 * Ms_ze_sx + C15 * Ms_zx_se + C55 * alpha
        out["M21s"] = C11 * zeta + C15 * 2.0 * Ms_ze_se + C55 * gamma
        out["M41s"] = C11 * epsil + C15 * 2.0 * Ms_zx_sx + C55 * beta
        out["M12s"] = C15 * delta + C13 * Ms_ze_sx + C55 * Ms_zx_se + C35 * alpha
        out["M22s"] = C15 * zeta + (C13 + C55) * Ms_ze_se + C35 * gamma
        out["M32s"] = C15 * delta + C13 * Ms_zx_se + C55 * Ms_ze_sx + C35 * alpha
        out["M42s"] = C15 * epsil + (C13 + C55) * Ms_zx_sx + C35 * beta
        out["M11z"] = C55 * delta + C35 * Ms_ze_sx + C35 * Ms_zx_se + C33 * alpha
        out["M21z"] = C55 * zeta + C35 * 2.0 * Ms_ze_se + C33 * gamma
        out["M41z"] = C55 * epsil + C35 * 2.0 * Ms_zx_sx + C33 * beta
        out["M1phi"] = C66 * delta + C46 * Ms_ze_sx + C46 * Ms_zx_se + C44 * alpha
        out["M2phi"] = C66 * zeta + C46 * 2.0 * Ms_ze_se + C44 * gamma
        out["M4phi"] = C66 * epsil + C46 * 2.0 * Ms_zx_sx + C44 * beta
        out["M_1"] = C12 * M_z_eta + C25 * M_s_eta
        out["M_2"] = C12 * M_z_xi + C25 * M_s_xi
        out["M_3"] = C23 * M_s_eta + C25 * M_z_eta
        out["M_4"] = C23 * M_s_xi + C25 * M_z_xi
        out["M_5"] = C66 * M_z_eta + C46 * M_s_eta
        out["M_6"] = C66 * M_z_xi + C46 * M_s_xi
        out["M_7"] = C44 * M_s_eta + C46 * M_z_eta
        out["M_8"] = C44 * M_s_xi + C46 * M_z_xi
        out["M_w1"] = (C22 + 4.0 * C66) * kw2
        out["M_w2"] = -2.0 * (C22 + C66) * kw2
        out["M_w3"] = 2.0 * C46 * kw2
        out["M_w4"] = (4.0 * C22 + C66) * kw2
        out["M_w5"] = 4.0 * C44 * kw2
        zero_axis(["M_1", "M_2", "M_3", "M_4", "M_5", "M_6", "M_7", "M_8",
                   "M_w1", "M_w2", "M_w3", "M_w4", "M_w5"])
        out["M0_w1"] = np.where(axm, (2.0 * a0(C12) + a0(C22) + 4.0 * a0(C66)) * ndf, 0.0)
        out["M0_w2"] = np.where(axm, -2.0 * (a0(C12) + a0(C22)) * ndf, 0.0)
        out["M0_w3"] = np.where(axm, (a0(C25) + 4.0 * a0(C46)) * ndf, 0.0)
        out["M0_w4"] = np.where(axm, (4.0 * a0(C22) - a0(C66)) * ndf, 0.0)
        out["M0_w5"] = np.where(axm, -2.0 * a0(C25) * ndf, 0.0)
        out["M0_w6"] = np.where(axm, 4.0 * a0(C44) * ndf, 0.0)
    else:
        raise ValueError(src_type)

    if anel:
        # def_precomp_terms.f90:1400-1416, 1481-1526.  In axial elements the reference
        # evaluates s at the *GLL* abscissa eta(ipol) instead of xi_k(ipol) (:1486);
        # reproduced here because these planes are inputs of the time loop.
        th_q = 0.5 * ((1.0 - basis.eta[None, :]) * es.th_a[:, None]
                      + (1.0 + basis.eta[None, :]) * es.th_b[:, None])
        s_q = g.r[:, :, None] * np.sin(th_q)[:, None, :]
        s_use = np.where(ax3, s_q, g.s)
        i0 = (np.arange(5)[None, None, :] == 0) & ax3
        out["Y"] = np.where(i0, 0.0, W2 * g.jac)
        out["V_s_eta"] = np.where(i0, 0.0, s_use * M_s_eta)
        out["V_s_xi"] = np.where(i0, 0.0, s_use * M_s_xi)
        out["V_z_eta"] = np.where(i0, 0.0, s_use * M_z_eta)
        out["V_z_xi"] = np.where(i0, 0.0, s_use * M_z_xi)
        out["Y0"] = np.where(axm, w0 * g.jac[:, :, 0], 0.0)
        out["V0_s_eta"] = np.zeros_like(ndf)
        out["V0_s_xi"] = np.where(axm, w0 * dsdxi0 * dsdxi0, 0.0)
        out["V0_z_eta"] = np.where(axm, w0 * dsdxi0 * g.dzdeta[:, :, 0], 0.0)
        out["V0_z_xi"] = np.where(axm, w0 * dsdxi0 * (-dzdxi0), 0.0)
        for n in ("Y", "V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi"):
            a = out[n]
            out[n + "_cg4"] = np.stack([a[:, 1, 1], a[:, 3, 1], a[:, 1, 3], a[:, 3, 3]], axis=1)
    return {k: _f32(v) for k, v in out.items()}


def cg4(a: np.ndarray) -> np.ndarray:
    """A(1,1), A(1,3), A(3,1), A(3,3) in Fortran (ipol,jpol) order
    (def_precomp_terms.f90:1581-1604) from a (nel, jpol, ipol) numpy array."""
    return np.stack([a[:, 1, 1], a[:, 3, 1], a[:, 1, 3], a[:, 3, 3]], axis=1)


def fluid_stiffness_terms(src_type: str, es: ElementSet, g: Geometry, rho,
                          basis: SpectralBasis) -> Dict[str, np.ndarray]:
    """def_precomp_terms.f90:2336-2470."""
    ij = 1.0 / g.jac
    W = g.W
    alpha = -ij * g.dsdxi * g.dsdeta * W
    beta = ij * g.dsdxi ** 2 * W
    gamma = ij * g.dsdeta ** 2 * W
    delta = -ij * g.dzdxi * g.dzdeta * W
    epsil = ij * g.dzdxi ** 2 * W
    zeta = ij * g.dzdeta ** 2 * W
    out = {"M1chi_fl": (delta + alpha) / rho,
           "M2chi_fl": (zeta + gamma) / rho,
           "M4chi_fl": (epsil + beta) / rho}
    ax = es.axis
    if src_type != "monopole":
        mw = g.massmat_kwts2 / rho
        mw = np.where(ax[:, None, None] & (np.arange(5)[None, None, :] == 0), 0.0, mw)
        m0 = np.where(ax[:, None], g.massmat_kwts2[:, :, 0] / rho[:, :, 0], 0.0)
        if src_type == "quadpole":
            mw = 4.0 * mw
            m0 = 4.0 * m0
        out["M_w_fl"] = mw
        out["M0_w_fl"] = m0
    return {k: _f32(v) for k, v in out.items()}


def pointwise_derivative_terms(es: ElementSet, g: Geometry) -> Dict[str, np.ndarray]:
    """def_precomp_terms.f90:178-225 (signs folded in, inv_s = 1 on the axis)."""
    with np.errstate(divide="ignore"):
        inv_s = np.where(g.s != 0.0, 1.0 / np.where(g.s != 0.0, g.s, 1.0), 1.0)
    inv_s = np.where(es.axis[:, None, None] & (np.arange(5)[None, None, :] == 0), 1.0, inv_s)
    return {"DsDeta_over_J": _f32(-g.dsdeta / g.jac), "DzDeta_over_J": _f32(g.dzdeta / g.jac),
            "DsDxi_over_J": _f32(g.dsdxi / g.jac), "DzDxi_over_J": _f32(-g.dzdxi / g.jac),
            "inv_s": _f32(inv_s)}


def sf_boundary_terms(mesh: LocalMesh, gs: Geometry, es: ElementSet) -> np.ndarray:
    """bdry_matr(0:4, nel_bdry, 2) — def_precomp_terms.f90:2503-2712 — returned as a
    numpy array of shape (2, nel_bdry, 5) (Fortran memory order).  `es`/`gs` describe the
    solid element of every boundary entry (same order as mesh.bdry_solid_el)."""
    b = mesh.basis
    nb = mesh.nel_bdry
    out = np.zeros((2, nb, 5))
    jj = mesh.bdry_jpol_solid
    k = np.arange(nb)
    th = gs.th                                   # (nb, 5)
    r = gs.r[k, jj]                              # (nb,)
    delta_th = 0.5 * np.abs(th[:, 4] - th[:, 0])
    ax = es.axis
    # non-axial
    out[0] = (delta_th[:, None] * b.wt[None, :]) * np.sin(th) * np.sin(th)
    out[1] = (delta_th[:, None] * b.wt[None, :]) * np.sin(th) * np.cos(th)
    if ax.any():
        w = b.wt_axial_k
        opx = 1.0 + b.xi_k
        a = np.nonzero(ax)[0]
        for i in range(1, 5):
            out[0, a, i] = delta_th[a] * w[i] * np.sin(th[a, i]) / opx[i] * np.sin(th[a, i])
            out[1, a, i] = delta_th[a] * w[i] * np.sin(th[a, i]) / opx[i] * np.cos(th[a, i])
        out[0, a, 0] = 0.0
        # note: the reference uses cos(0)=1 here also at the southern axis (:2603)
        out[1, a, 0] = 1.0 / r[a] * delta_th[a] * w[0] * gs.dsdxi[a, jj[a], 0]
    sign = np.where(mesh.bdry_above, 1.0, -1.0)
    out *= (sign * r * r)[None, :, None]
    return _f32(out)


# --------------------------------------------------------------------------------------
@dataclass
class AttenuationModel:
    """What `prepare_attenuation` (attenuation.f90:682-1095) leaves behind for the loop.
    The SLS fit (a random search with an unseeded RNG, :1183-1339) is not reproducible in the
    reference; `w_j`, `y_j` are therefore inputs here.  Defaults: a rounded log-spaced 5-SLS set
    for 1 mHz - 1 Hz whose Q is flat to 16 % (tests/test_sls_fit.py); `invert_linear_solids` below
    is the reference's search with a seed (flat to 2 % over the upper two decades) and
    `AttenuationModel.fitted(...)` the set it finds."""
    n_sls: int = 5
    w_j: np.ndarray = field(default_factory=lambda: 2 * np.pi * np.array(
        [0.0015, 0.0090, 0.052, 0.29, 1.55]))
    y_j: np.ndarray = field(default_factory=lambda: np.array(
        [1.53, 1.16, 1.21, 1.09, 1.70]))
    f_min: float = 0.001
    f_max: float = 1.0
    w_0: float = 1.0            # reference frequency of the background model [Hz]
    do_corr_lowq: bool = True
    coarse_grained: bool = True

    @classmethod
    def fitted(cls, n_sls: int = 5, f_min: float = 0.001, f_max: float = 1.0, seed: int = 0, max_it: int = 100000, **kw):
        """The set the reference's search finds for this band (NR_LIN_SOLIDS, F_MIN, F_MAX, MAXINT_SA of inparam_advanced)."""
        w_j, y_j, _ = invert_linear_solids(n_sls, f_min, f_max, seed=seed, max_it=max_it)
        return cls(n_sls=n_sls, w_j=w_j, y_j=y_j, f_min=f_min, f_max=f_max, **kw)


def q_linear_solid(y_j, w_j, w, exact: bool = False) -> np.ndarray:
    """Q(w) of a set of standard linear solids (Emmerich & Korn eq. 21 inverted, or its linearisation eq. 22;
    attenuation.f90:1099-1130)."""
    y_j, w_j, w = (np.asarray(a, dtype=np.float64) for a in (y_j, w_j, w))
    num = np.ones_like(w)
    if exact:
        num = num + (y_j[:, None] * w[None, :] ** 2 / (w[None, :] ** 2 + w_j[:, None] ** 2)).sum(axis=0)
    den = (y_j[:, None] * w[None, :] * w_j[:, None] / (w[None, :] ** 2 + w_j[:, None] ** 2)).sum(axis=0)
    return num / den


def invert_linear_solids(n_sls: int = 5, f_min: float = 0.001, f_max: float = 1.0, *, Q: float = 1.0, nfsamp: int = 100,
                         max_it: int = 100000, Tw: float = 0.1, Ty: float = 0.1, d: float = 0.99995, fixfreq: bool = False,
                         freq_weight: bool = True, w_ref: float = 1.0, alpha: float = 0.0, exact: bool = False, seed: int = 0):
    """The reference's fit of the SLS set to a constant (or power-law) Q: a random search that perturbs the
    relaxation frequencies and amplitudes within a shrinking range and keeps what lowers the frequency-weighted
    log-l2 misfit (invert_linear_solids + l2_error, attenuation.f90:1159-1339; the defaults are those of
    inparam_advanced: NR_F_SAMPLE 100, MAXINT_SA 100000, TSTART_SR / TSTART_AMP 0.1, T_DECAY 0.99995, FREQ_WEIGHT
    true).  The reference draws from an unseeded random_number; here the stream is seeded, so a fit can be
    repeated.  Returns (w_j [rad/s], y_j for Q = 1 — the loop divides by Q —, misfit per iteration)."""
    rng = np.random.default_rng(seed)
    if n_sls > 1:
        expo = (np.log10(f_max) - np.log10(f_min)) / (n_sls - 1.0)
        w_j = 2 * np.pi * 10.0 ** (np.log10(f_min) + np.arange(n_sls) * expo)
    else:
        w_j = np.array([np.sqrt(f_max * f_min) * 2 * np.pi])
    expo = (np.log10(f_max) - np.log10(f_min)) / (nfsamp - 1.0)
    w = 2 * np.pi * 10.0 ** (np.log10(f_min) + np.arange(nfsamp) * expo)
    q_target = Q * (w / w_ref) ** alpha
    weights = w / w.sum() * nfsamp if freq_weight else np.ones(nfsamp)

    def misfit(y, wj):
        return np.sqrt((np.log(q_target / q_linear_solid(y, wj, w, exact)) ** 2 * weights).sum() / float(nfsamp))

    y_j = np.full(n_sls, 1.0 / Q * 1.5)
    chi = misfit(y_j, w_j)
    chil = np.zeros(max_it)
    r = rng.random((max_it, n_sls, 2))
    for it in range(max_it):
        w_t = w_j if fixfreq else w_j * (1.0 + (0.5 - r[it, :, 0]) * Tw)
        y_t = y_j * (1.0 + (0.5 - r[it, :, 1]) * Ty)
        c = misfit(y_t, w_t)
        Tw *= d
        Ty *= d
        if c < chi:
            y_j, w_j, chi = y_t, w_t, c
        chil[it] = chi
    return w_j, y_j, chil


def fast_correct(y_j: np.ndarray) -> np.ndarray:
    """attenuation.f90:1139-1155."""
    dy = np.zeros_like(y_j)
    dy[0] = 1.0 + 0.5 * y_j[0]
    for k in range(1, y_j.size):
        dy[k] = dy[k - 1] + (dy[k - 1] - 0.5) * y_j[k - 1] + 0.5 * y_j[k]
    return y_j * dy


def attenuation_terms(att: AttenuationModel, deltat: float, es: ElementSet, g: Geometry,
                      lam, mu, qmu, qka):
    """Time-step factors, delta moduli and unrelaxed moduli (attenuation.f90:882-1063).
    Returns (dict of loop inputs, lam_unrelaxed, mu_unrelaxed)."""
    w_j = np.asarray(att.w_j, dtype=np.float64)
    y_j = np.asarray(att.y_j, dtype=np.float64)
    exp_w = np.exp(-w_j * deltat)
    ts_fac_tm1 = (1.0 - exp_w) / (w_j * deltat) - exp_w
    ts_fac_t = (exp_w - 1.0) / (w_j * deltat) + 1.0
    w_0 = att.w_0 * 2 * np.pi
    w_1 = np.sqrt(att.f_min * att.f_max) * 2 * np.pi
    nel = es.nel
    # coarse-grained weights (attenuation.f90:940-996)
    gw = g.massmat_k.copy()          # = wt wt J s (non-axial) / axial analogue incl. axis
    # the reference's axial gamma_w_l(i>0) = wax w/(1+xi) J s ; i=0: wax w J dsdxi
    # both equal massmat_k = J * s/(1+xi) * wax * w, so massmat_k is reused.
    G = lambda i, j: gw[:, j, i]
    wcg = np.zeros((nel, 5, 5))
    wcg[:, 1, 1] = (G(0, 0) + G(0, 1) + G(1, 0) + G(1, 1)
                    + 0.5 * (G(0, 2) + G(1, 2) + G(2, 0) + G(2, 1)) + 0.25 * G(2, 2)) / G(1, 1)
    wcg[:, 3, 1] = (G(0, 3) + G(0, 4) + G(1, 3) + G(1, 4)
                    + 0.5 * (G(0, 2) + G(1, 2) + G(2, 3) + G(2, 4)) + 0.25 * G(2, 2)) / G(1, 3)
    wcg[:, 1, 3] = (G(3, 0) + G(3, 1) + G(4, 0) + G(4, 1)
                    + 0.5 * (G(2, 0) + G(2, 1) + G(3, 2) + G(4, 2)) + 0.25 * G(2, 2)) / G(3, 1)
    wcg[:, 3, 3] = (G(3, 3) + G(3, 4) + G(4, 3) + G(4, 4)
                    + 0.5 * (G(2, 3) + G(2, 4) + G(3, 2) + G(4, 2)) + 0.25 * G(2, 2)) / G(3, 3)

    mu_fac = np.zeros(nel)
    ka_fac = np.zeros(nel)
    sum_mu = np.zeros(nel)
    sum_ka = np.zeros(nel)
    for q in np.unique(qmu):
        yp = fast_correct(y_j / q) if att.do_corr_lowq else y_j / q
        m = qmu == q
        mu_fac[m] = np.sum(yp * w_j ** 2 / (w_1 ** 2 + w_j ** 2)) / yp.sum()
        sum_mu[m] = yp.sum()
    for q in np.unique(qka):
        yp = fast_correct(y_j / q) if att.do_corr_lowq else y_j / q
        m = qka == q
        ka_fac[m] = np.sum(yp * w_j ** 2 / (w_1 ** 2 + w_j ** 2)) / yp.sum()
        sum_ka[m] = yp.sum()
    e3 = lambda a: a[:, None, None]
    mu_w1 = mu * (1.0 + 2.0 / (np.pi * e3(qmu)) * np.log(w_1 / w_0))
    ka_w1 = (lam + 2.0 / 3.0 * mu) * (1.0 + 2.0 / (np.pi * e3(qka)) * np.log(w_1 / w_0))
    dmu0 = mu_w1 / e3(1.0 / sum_mu + 1.0 - mu_fac)
    dka0 = ka_w1 / e3(1.0 / sum_ka + 1.0 - ka_fac)
    out = {"exp_w_j_deltat": exp_w, "ts_fac_t": ts_fac_t, "ts_fac_tm1": ts_fac_tm1,
           "y_j": y_j, "w_j": w_j}
    if att.coarse_grained:
        mu_u = mu_w1 + wcg * dmu0 * e3(mu_fac)
        lam_u = ka_w1 + wcg * dka0 * e3(ka_fac) - 2.0 / 3.0 * mu_u
        out["delta_mu_cg4"] = _f32(cg4(wcg * dmu0))
        out["delta_kappa_cg4"] = _f32(cg4(wcg * dka0))
    else:
        mu_u = mu_w1 + dmu0 * e3(mu_fac)
        lam_u = ka_w1 + dka0 * e3(ka_fac) - 2.0 / 3.0 * mu_u
        out["delta_mu"] = _f32(dmu0)
        out["delta_kappa"] = _f32(dka0)
    out["Q_mu"] = _f32(qmu)
    out["Q_kappa"] = _f32(qka)
    return out, lam_u, mu_u
