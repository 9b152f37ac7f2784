"""Source time function and source-term arrays (host side; inputs of the time loop).

Restates SOLVER/source.f90: `gauss`/`gauss_d`/`gauss_dd`/`errorf` (:587-660) and their
point-wise twins used by the symplectic schemes (`compute_stf_t`, :206-233),
`define_bodyforce` (:921-978), `define_moment_tensor` (:985-1226), `compute_src`
(:237-431).  The device only ever receives `stf(niter)`, `ielsrc(<=8)` and
`source_term_el(0:4,0:4,8,3)`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np

from .mesh import LocalMesh

SRC_POLE = {
    "explosion": "monopole", "mrr": "monopole", "mtt_p_mpp": "monopole", "vertforce": "monopole",
    "mtr": "dipole", "mpr": "dipole", "thetaforce": "dipole", "phiforce": "dipole",
    "mtp": "quadpole", "mtt_m_mpp": "quadpole",
}


@dataclass
class SourceParams:
    src_type2: str = "explosion"       # src_type(2) in the reference
    depth: float = 100.0e3             # metres below the surface, on the northern axis
    magnitude: float = 1.0e20
    stf_type: str = "gauss_0"
    t_0: float = 50.0                  # dominant period [s]
    decay: float = 3.5
    shift_fact: float = 1.5            # in units of t_0 (rounded to a multiple of deltat)

    @property
    def src_type1(self) -> str:
        return SRC_POLE[self.src_type2]


def stf_at(p: SourceParams, t: np.ndarray, deltat: float) -> np.ndarray:
    """Source time function at times t (float64) — gauss_t etc., source.f90:206-233."""
    shift = np.ceil(p.shift_fact * p.t_0 / deltat) * deltat
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
    raise ValueError(p.stf_type)


def compute_stf(p: SourceParams, niter: int, deltat: float) -> np.ndarray:
    """stf(1:niter) in single precision; t = i*deltat is rounded to realkind first, as
    in source.f90:590-593."""
    t = (np.arange(1, niter + 1, dtype=np.float64) * deltat).astype(np.float32).astype(np.float64)
    return stf_at(p, t, deltat).astype(np.float32)


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
