"""Source time function and source-term arrays (host side; inputs of the time loop).

Restates SOLVER/source.f90: `compute_stf` (:144-202) with `gauss`/`gauss_d`/`gauss_dd`/`errorf`
(:587-660), the Numerical-Recipes `erf` (:662-692) and the discrete Dirac / quasi-Heaviside of
`delta_src` (:696-814); their point-wise twins used by the symplectic schemes (`compute_stf_t`,
:206-233, :818-917); the choice of the discrete Dirac and of the source shift in
`parameters.F90:975-1072`; `define_bodyforce` (:921-978), `define_moment_tensor` (:985-1226),
`compute_src` (:237-431).  The device only ever receives `stf(niter)`, `ielsrc(<=8)` and
`source_term_el(0:4,0:4,8,3)`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from .mesh import LocalMesh

SRC_POLE = {
    "explosion": "monopole", "mrr": "monopole", "mtt_p_mpp": "monopole", "vertforce": "monopole",
    "mtr": "dipole", "mpr": "dipole", "thetaforce": "dipole", "phiforce": "dipole",
    "mtp": "quadpole", "mtt_m_mpp": "quadpole",
}

# SOURCE_FUNCTION values of inparam_advanced (compute_stf, source.f90:152-171)
STF_NAMES = ("gauss_0", "gauss_1", "gauss_2", "errorf", "dirac_0", "dirac_1", "quheavi")
DIRAC_APPROX = ("cauchy", "caulor", "sincfc", "gaussi", "triang", "1dirac")   # source.f90:714
_PI = 3.1415926535898            # global_parameters.f90


@dataclass
class SourceParams:
    src_type2: str = "explosion"       # src_type(2) in the reference
    depth: float = 100.0e3             # metres below the surface, on the northern axis
    magnitude: float = 1.0e20
    stf_type: str = "gauss_0"
    t_0: float = 50.0                  # dominant period [s]; dirac_0 / quheavi: discrete_dirac_halfwidth
    decay: float = 3.5
    shift_fact: float = 1.5            # in units of t_0 (rounded to a multiple of deltat)
    shift_seconds: Optional[float] = None   # the shift in seconds where the caller fixed it
                                            # (discrete_dirac_setup); None: from shift_fact
    discrete_choice: str = "gaussi"    # delta_src's approximation of the Dirac (parameters.F90:995-999)

    @property
    def src_type1(self) -> str:
        return SRC_POLE[self.src_type2]


def stf_shift(p: SourceParams, deltat: float) -> float:
    """shift_fact of the reference in seconds."""
    fixed = getattr(p, "shift_seconds", None)      # containers written before the field existed
    if fixed is not None:
        return float(fixed)
    return float(np.ceil(p.shift_fact * p.t_0 / deltat) * deltat)


def discrete_dirac_setup(period: float, deltat: float, seis_it: int, strain_it: int = 0,
                         dump_wavefields: bool = False, stf_type: str = "dirac_0") -> Tuple[float, str, float]:
    """(discrete_dirac_halfwidth, discrete_choice, shift_fact_discrete_dirac) for dirac_0 / quheavi as
    parameters.F90:975-1068 sets them: half width period / 8 (period / (2 deltat_coarse), at least
    15, where only seismograms are down-sampled), a Gaussian where anything is down-sampled and a
    one-sample spike otherwise, and the first shift beyond four half widths that is a whole number of
    time steps, seismogram samples and wavefield samples.  (The narrower half width is taken for
    dirac_0 only: the reference tests stf_type against 'queavi' there, parameters.F90:979.)"""
    seis_dt = deltat * seis_it
    deltat_coarse = deltat * (strain_it if dump_wavefields and strain_it > 0 else seis_it)
    pvh = 8
    if not dump_wavefields and stf_type == "dirac_0" and deltat_coarse > 1.9 * deltat:
        pvh = int(period / (2.0 * deltat_coarse))           # integer period_vs_discrete_halfwidth
        if pvh < 15:
            pvh = 15
    choice = "gaussi" if (dump_wavefields or deltat_coarse > 1.9 * deltat) else "1dirac"
    half = float(np.float32(period / pvh))                  # realkind
    n = int(np.ceil(4.0 * half / deltat))
    for i in range(1, n + 1):
        dshift = deltat * n + float(i) * deltat
        if (abs(round(dshift / deltat_coarse) - dshift / deltat_coarse) < 0.01 * deltat
                and abs(round(dshift / deltat) - dshift / deltat) < 0.01 * deltat
                and abs(round(dshift / seis_dt) - dshift / seis_dt) < 0.01 * deltat):
            return half, choice, float(np.float32(deltat_coarse * np.ceil(dshift / deltat_coarse)))
    raise ValueError("no source shift is a multiple of deltat, seis_dt and deltat_coarse "
                     "(the reference stops: 'source time shift not defined')")


_ERF_COEFFS = np.array([-1.26551223, 1.00002368, 0.37409196, 0.09678418, -0.18628806,
                        0.27886807, -1.13520398, 1.48851587, -0.82215223, 0.17087277],
                       dtype=np.float32).astype(np.float64)   # default-real literals (source.f90:673-675)


def erf_nr(x: np.ndarray) -> np.ndarray:
    """The reference's own error function (Numerical Recipes erfc, source.f90:662-692)."""
    x = np.asarray(x, dtype=np.float64)
    z = np.abs(x)
    t = 1.0 / (1.0 + 0.5 * z)
    poly = np.full_like(z, _ERF_COEFFS[-1])
    for c in _ERF_COEFFS[-2::-1]:
        poly = t * poly + c
    erfcc = t * np.exp(-z * z + poly)
    erfcc = np.where(x < 0.0, 2.0 - erfcc, erfcc)
    return 1.0 - erfcc


def stf_at(p: SourceParams, t: np.ndarray, deltat: float) -> np.ndarray:
    """The smooth source time functions at times t (float64) — gauss_t, gauss_d_t, gauss_dd_t,
    errorf_t (source.f90:818-886)."""
    shift = stf_shift(p, deltat)
    a = p.decay / p.t_0
    x = a * (t - shift)
    if p.stf_type == "gauss_0":
        return np.exp(-x ** 2) * p.magnitude * a / np.sqrt(np.pi)
    if p.stf_type == "gauss_1":
        return (-2.0 * a ** 2 * (t - shift) * np.exp(-x ** 2)
                / (a * np.sqrt(2.0) * np.exp(-0.5)) * p.magnitude)
    if p.stf_type == "gauss_2":
        return (a ** 2 * (2.0 * a ** 2 * (t - shift) ** 2 - 1.0) * np.exp(-x ** 2)
                / (2.0 * a ** 2 * np.exp(-1.5)) * p.magnitude)
    if p.stf_type == "errorf":
        return (erf_nr(x) * 0.5 + 0.5) * p.magnitude
    raise ValueError(p.stf_type)


def delta_src(p: SourceParams, niter: int, deltat: float) -> np.ndarray:
    """delta_src (source.f90:696-814): the discrete Dirac chosen by `discrete_choice`, normalised to
    unit integral, times the magnitude; for quheavi its running integral.  float32 like `stf`."""
    if p.discrete_choice not in DIRAC_APPROX:
        raise ValueError(f"do not know discrete Dirac {p.discrete_choice}")
    a = float(np.float32(p.t_0))                       # discrete_dirac_halfwidth (realkind)
    shift = float(np.float32(stf_shift(p, deltat)))    # shift_fact_discrete_dirac (realkind)
    i = np.arange(1, niter + 1)
    t = i.astype(np.float64) * deltat
    c = p.discrete_choice
    if c == "cauchy":
        signal = 1.0 / a * np.exp(-np.abs((t - shift) / a))
    elif c == "caulor":
        signal = 1.0 / _PI * a / (a ** 2 + (t - shift) ** 2)
    elif c == "sincfc":
        t = np.where(t == shift, 0.00001 + shift, t)
        signal = 1.0 / (a * _PI) * np.sin((-shift + t) / a) / ((-shift + t) / a)
    elif c == "gaussi":
        signal = 1.0 / (a * np.sqrt(_PI)) * np.exp(-((t - shift) / a) ** 2)
    elif c == "triang":
        signal = np.where(np.abs(t - shift) <= a / 2.0, 2.0 / a - 4.0 / a ** 2 * np.abs(t - shift), 0.0)
    else:                                              # 1dirac: one non-zero sample
        signal = np.where(i == int(shift / deltat), 1.0, 0.0)
    integral = signal.sum() * deltat
    stf = (signal / integral).astype(np.float32)
    stf = (stf.astype(np.float64) * p.magnitude).astype(np.float32)
    if p.stf_type == "quheavi":                        # int_stf(i) = int_stf(i-1) + stf(i) deltat, in dp
        return np.cumsum(stf.astype(np.float64) * deltat).astype(np.float32)
    return stf


def compute_stf(p: SourceParams, niter: int, deltat: float) -> np.ndarray:
    """stf(1:niter) in single precision (compute_stf, source.f90:144-202); for the smooth types
    t = i*deltat is rounded to realkind first, as in source.f90:590-593."""
    if p.stf_type in ("dirac_0", "dirac_1", "quheavi"):
        return delta_src(p, niter, deltat)
    t = (np.arange(1, niter + 1, dtype=np.float64) * deltat).astype(np.float32).astype(np.float64)
    return stf_at(p, t, deltat).astype(np.float32)


def compute_stf_t(p: SourceParams, subdt: np.ndarray, deltat: float, seis_it: int = 1) -> np.ndarray:
    """compute_stf_t (source.f90:208-233) on the sub-stage times of one symplectic step (float64).
    dirac_0 is the hat function of delta_src_t (:890-904), switched by the *first* sub-stage time of
    the step; quheavi is quasiheavi_t (:908-917), which sets the sub-stages seis_it..nstages — an
    index, not a time — to the magnitude."""
    subdt = np.asarray(subdt, dtype=np.float64)
    if p.stf_type == "dirac_0":
        shift = stf_shift(p, deltat)
        out = np.zeros_like(subdt)
        if subdt[0] > shift - deltat and subdt[0] <= shift:
            out = (subdt - subdt[0]) / deltat * p.magnitude / deltat
        if subdt[0] >= shift and subdt[0] < shift + deltat:
            out = (1.0 - (subdt - subdt[0]) / deltat) * p.magnitude / deltat
        return out
    if p.stf_type == "quheavi":
        out = np.zeros_like(subdt)
        out[max(seis_it, 1) - 1:] = p.magnitude
        return out
    if p.stf_type == "dirac_1":
        raise ValueError("source time function non existant for the symplectic schemes: dirac_1")
    return stf_at(p, subdt, deltat)


def _mxm(a, b):
    """Fortran mxm(a,b)(i,j) = sum_k a(i,k) b(k,j) on numpy arrays indexed [j,i]."""
    # a_f(i,k) = a[k,i]; result_f(i,j) = sum_k a[k,i]*b[j,k] -> res[j,i]
    return np.einsum("ki,jk->ji", a, b)


def compute_source_terms(mesh: LocalMesh, p: SourceParams, pw: Dict[str, np.ndarray]):
    """(nelsrc, ielsrc[8] 1-based, source_term_el[3,8,5,5]) for a point source on the
    northern axis.  `pw` are the solid pointwise-derivative planes of this rank."""
    src1 = p.src_type1
    st = np.zeros((3, 8, 5, 5), dtype=np.float32)
    ielsrc = np.zeros(8, dtype=np.int32)
    spec, es, b = mesh.spec, mesh.solid, mesh.basis
    zsrc = spec.router - p.depth
    # find_srcloc (source.f90:454-476): on-axis GLL point closest in z, over all ranks
    cand = np.nonzero(es.axis & es.north)[0]
    eta = b.eta
    if mesh.it0 != 0 or cand.size == 0:
        return 0, ielsrc, st
    r = 0.5 * ((1 - eta)[None, :] * es.r_a[cand, None] + (1 + eta)[None, :] * es.r_b[cand, None])
    d = np.abs(r - zsrc)
    dmin = d.min()
    sol = np.nonzero(~spec.fluid_ir)[0]
    r_all = 0.5 * ((1 - eta)[None, :] * spec.r_edges[sol, None] + (1 + eta)[None, :] * spec.r_edges[sol + 1, None])
    if dmin > np.abs(r_all - zsrc).min() * (1 + 1e-12) + 1e-6:
        return 0, ielsrc, st                                   # another radial block holds the source
    jh = np.argwhere(d <= dmin * (1 + 1e-12) + 1e-6)
    for a, j in jh:
        e = cand[a]
        on_cut = (es.ir[e] == mesh.ir0 and j == 0 and mesh.ir0 > 0) or (es.ir[e] == mesh.ir1 - 1 and j == 4 and mesh.ir1 < spec.nr)
        assert not (mesh.nranks_r > 1 and on_cut), "source on a radial cut: move the source or the cut"
    hits = np.argwhere(d <= dmin * (1 + 1e-12) + 1e-6)
    srcs = [(cand[a], 0, j) for a, j in hits][:2]          # (iel, ipol, jpol)
    # work only on the elements that share a global point with a source element
    ig_all = mesh.igloc_solid.astype(np.int64).reshape(es.nel, 25)
    src_gids = np.unique(np.concatenate([ig_all[e] for (e, _, _) in srcs]))
    cand_el = np.nonzero(np.isin(ig_all, src_gids).any(axis=1))[0]
    loc = {int(e): k for k, e in enumerate(cand_el)}
    ncand = cand_el.size
    source_term = np.zeros((ncand, 3, 5, 5), dtype=np.float64)
    G2 = b.G2.astype(np.float64)
    G2T = b.G2T.astype(np.float64)
    G1T = b.G1T.astype(np.float64)
    nsrc = len(srcs)
    force = p.src_type2 in ("vertforce", "thetaforce", "phiforce")
    for (e, ip, jp) in srcs:
        q = loc[int(e)]
        if force:
            # define_bodyforce (source.f90:921-943): unit value at the source point
            source_term[q, 2 if p.src_type2 == "vertforce" else 0, jp, ip] = 1.0
            continue
        # pointwise planes are numpy [j,i]; transpose to Fortran (i,j)
        dzdeta = pw["DzDeta_over_J"][e].astype(np.float64).T
        dzdxi = pw["DzDxi_over_J"][e].astype(np.float64).T
        dsdeta = pw["DsDeta_over_J"][e].astype(np.float64).T
        dsdxi = pw["DsDxi_over_J"][e].astype(np.float64).T
        GT = G1T if es.axis[e] else G2T
        for ipol in range(5):
            for jpol in range(5):
                ws = np.zeros((5, 5))          # Fortran ws(i,j)
                ws[ipol, jpol] = 1.0
                mxm1 = GT @ ws                  # sum_k GT(i,k) ws(k,j)
                mxm2 = ws @ G2                  # sum_k ws(i,k) G2(k,j)
                dsws = dzdeta * mxm1 + dzdxi * mxm2      # dsdf_elem_solid
                dzwz = dsdeta * mxm1 + dsdxi * mxm2      # dzdf_elem_solid
                if src1 == "monopole":
                    if p.src_type2 == "explosion":
                        source_term[q, 0, jpol, ipol] = 2.0 * dsws[ip, jp]
                        source_term[q, 2, jpol, ipol] = dzwz[ip, jp]
                    elif p.src_type2 == "mtt_p_mpp":
                        source_term[q, 0, jpol, ipol] = dsws[ip, jp]
                    elif p.src_type2 == "mrr":
                        source_term[q, 2, jpol, ipol] = dzwz[ip, jp]
                    else:
                        raise ValueError(p.src_type2)
                elif src1 == "dipole":
                    source_term[q, 0, jpol, ipol] = dzwz[ip, jp]
                    source_term[q, 2, jpol, ipol] = dsws[ip, jp]
                else:
                    source_term[q, 0, jpol, ipol] = dsws[ip, jp]
                    source_term[q, 1, jpol, ipol] = dsws[ip, jp]
    if not force:
        source_term /= float(nsrc)
    # assembly over the local mesh (pdistsum_solid, source.f90:1133; bodyforce :940)
    ig = ig_all[cand_el].reshape(-1)
    uniq, inv = np.unique(ig, return_inverse=True)
    for c in range(3):
        flat = source_term[:, c].reshape(-1)
        g = np.zeros(uniq.size)
        np.add.at(g, inv, flat)
        source_term[:, c] = g[inv].reshape(ncand, 5, 5)
    if not force:
        source_term[np.abs(source_term) < 1e-30] = 0.0       # smallval cut (:1136-1147)
    # normalisation (source.f90:273-330)
    if src1 == "dipole":
        source_term /= np.pi
    else:
        source_term /= 2.0 * np.pi
    # axis masks (source.f90:1160-1170); single forces are not masked in the reference
    if not force:
        axm = es.axis[cand_el]
        if src1 == "monopole":
            source_term[axm, 0, :, 0] = 0.0
        elif src1 == "dipole":
            source_term[axm, 1, :, 0] = 0.0
            source_term[axm, 2, :, 0] = 0.0
        else:
            source_term[axm, :, :, 0] = 0.0
    k = 0
    for q in range(ncand):
        if np.abs(source_term[q]).max() > 0:
            assert k < 8, "more than 8 source elements"
            ielsrc[k] = cand_el[q] + 1
            st[:, k] = source_term[q].astype(np.float32)
            k += 1
    return k, ielsrc, st
