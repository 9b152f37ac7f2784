"""Assemble one rank's complete set of time-loop inputs (what `prepare_waves`,
SOLVER/time_evol_wave.F90:47-225, leaves in the Fortran modules) for a synthetic mesh.

`build_problem` is the stand-in for the Fortran host in this repository: it produces
exactly the arrays the C ABI (`include/axisem_b200.h`) takes, in the reference's layouts
and index conventions.  The same `Problem` feeds the CUDA library, the CPU oracle and the
benchmark, so parity tests compare like with like.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .mesh import (ElementSet, LocalMesh, MeshSpec, build_rank, element_coords,
                   make_elements, surface_receivers)
from .model import evaluate_layer
from .precomp import (AttenuationModel, attenuation_terms, fluid_stiffness_terms, geometry,
                      material, pointwise_derivative_terms, sf_boundary_terms,
                      solid_stiffness_terms, cg4, _f32)
from .source import SourceParams, compute_source_terms, compute_stf
from .spectral import SpectralBasis

SRC_ORDER = {"monopole": 0, "dipole": 1, "quadpole": 2}
TIME_SCHEMES = ("newmark2", "symplec4", "ML_SO4m5", "ML_SO6m7", "KL_O8m17", "SS_35o10")


def stable_timestep(spec: MeshSpec, basis: SpectralBasis, courant: float = 0.6) -> float:
    """Global time step from the smallest GLL spacing / fastest P velocity, evaluated on
    the spec alone so that every rank gets the same number (the reference reads dt from
    the mesher; MANUAL/theoretical_foundations.tex:319-341, Courant 0.6)."""
    frac = 0.5 * (basis.eta[1] - basis.eta[0])
    dth = np.pi / spec.ntheta
    dt = np.inf
    for ir in range(spec.nr):
        L = spec.layers[spec.layer_of_ir[ir]]
        r0, r1 = spec.r_edges[ir], spec.r_edges[ir + 1]
        vp = max(evaluate_layer(L, np.array([r0, r1]))[6])
        h = min((r1 - r0), r0 * dth) * frac
        dt = min(dt, courant * h / vp)
    return float(dt)


@dataclass
class Problem:
    mesh: LocalMesh
    src_type: str
    src_order: int
    time_scheme: str
    deltat: float
    niter: int
    seis_it: int
    strain_it: int
    anel: bool
    solid: Dict[str, np.ndarray]
    fluid: Dict[str, np.ndarray]
    inv_mass_rho: np.ndarray
    inv_mass_fluid: np.ndarray
    fluid_free_surface_mask: np.ndarray
    inv_rho_fluid: np.ndarray
    pw_solid: Dict[str, np.ndarray]
    pw_fluid: Dict[str, np.ndarray]
    bdry_matr: np.ndarray
    att: Optional[Dict[str, np.ndarray]]
    att_model: Optional[AttenuationModel]
    source: SourceParams
    nelsrc: int
    ielsrc: np.ndarray
    source_term_el: np.ndarray
    stf: np.ndarray
    recfile_el: np.ndarray
    rec_index: np.ndarray
    solid_absorbing_gamma: Optional[np.ndarray] = None
    fluid_absorbing_gamma: Optional[np.ndarray] = None
    kwf: Optional[Dict[str, np.ndarray]] = None

    @property
    def num_rec(self):
        return int(self.recfile_el.shape[0])


def _assembled_mass(spec: MeshSpec, basis: SpectralBasis, own: ElementSet, it0: int, it1: int,
                    fluid: bool, own_val: np.ndarray, valfun) -> np.ndarray:
    """Direct-stiffness-sum a per-point quantity over the *whole* mesh and return it on
    this rank's elements (def_mass_matrix_k calls pdistsum_* once, :756, :820).  Ghost
    columns of the neighbouring slices supply the cross-rank contributions."""
    nrn = spec.nrnode_fluid if fluid else spec.nrnode_solid
    if own.nel == 0:
        return own_val
    lo = max(it0 - 1, 0)
    hi = min(it1 + 1, spec.ntheta)
    ghost_cols = [c for c in (it0 - 1, it1) if 0 <= c < spec.ntheta and not (it0 <= c < it1)]
    sets = [(own, own_val)]
    for c in ghost_cols:
        irs = np.nonzero(spec.fluid_ir == fluid)[0]
        es = make_elements(spec, np.full(irs.size, c), irs)
        sets.append((es, valfun(es)))
    dense = np.zeros(((hi - lo) * 4 + 1, nrn), dtype=np.float32)
    idx = []
    i = np.arange(5)
    for es, val in sets:
        ip = np.where(es.north[:, None], i[None, :], 4 - i[None, :])
        tn = 4 * (es.it[:, None] - lo) + ip
        rn = spec.rbase[es.ir][:, None] + ip
        T = np.broadcast_to(tn[:, None, :], val.shape)
        Rn = np.broadcast_to(rn[:, :, None], val.shape)
        np.add.at(dense, (T.reshape(-1), Rn.reshape(-1)), val.astype(np.float32).reshape(-1))
        idx.append((T, Rn))
    T, Rn = idx[0]
    return dense[T, Rn]


def build_problem(spec: MeshSpec, source: Optional[SourceParams] = None, *, rank: int = 0,
                  nranks: int = 1, anel: bool = False, att: Optional[AttenuationModel] = None,
                  time_scheme: str = "newmark2", niter: int = 100, seis_it: int = 1,
                  strain_it: int = 0, courant: float = 0.6, deltat: Optional[float] = None,
                  rec_colat_deg=None, dump: bool = False) -> Problem:
    assert time_scheme in TIME_SCHEMES
    source = source or SourceParams()
    src_type = source.src_type1
    basis = SpectralBasis(spec.npol)
    mesh = build_rank(spec, rank, nranks, basis)
    if deltat is None:
        deltat = stable_timestep(spec, basis, courant)
        if time_scheme != "newmark2":
            deltat *= 1.5                      # time_evol_wave.F90:511
    gs = geometry(mesh.solid, basis)
    gf = geometry(mesh.fluid, basis)
    rho_s, lam_s, mu_s, xi_s, phi_s, eta_s, _, qmu, qka = material(spec, mesh.solid, gs)
    rho_f, lam_f, mu_f, *_ = material(spec, mesh.fluid, gf)

    # ---- mass matrices (before attenuation changes the moduli, as in the reference) ---
    def rho_mass(es):
        g = geometry(es, basis)
        rho = material(spec, es, g)[0]
        return rho * g.massmat_k

    def lam_mass(es):
        g = geometry(es, basis)
        lam = material(spec, es, g)[1]
        return g.massmat_k / lam

    m_s = _assembled_mass(spec, basis, mesh.solid, mesh.it0, mesh.it1, False,
                          rho_s * gs.massmat_k, rho_mass)
    inv_mass_rho = (1.0 / m_s.astype(np.float64)) if mesh.nel_solid else m_s
    if src_type == "dipole":
        inv_mass_rho = 0.5 * inv_mass_rho                 # def_precomp_terms.f90:773
    if mesh.nel_fluid:
        m_f = _assembled_mass(spec, basis, mesh.fluid, mesh.it0, mesh.it1, True,
                              gf.massmat_k / lam_f, lam_mass)
        inv_mass_fluid = 1.0 / m_f.astype(np.float64)
        inv_rho_fluid = 1.0 / rho_f
    else:
        inv_mass_fluid = np.zeros((0, 5, 5))
        inv_rho_fluid = np.zeros((0, 5, 5))

    pw_s = pointwise_derivative_terms(mesh.solid, gs)
    pw_f = pointwise_derivative_terms(mesh.fluid, gf)

    att_d = None
    if anel:
        att = att or AttenuationModel()
        att_d, lam_s, mu_s = attenuation_terms(att, deltat, mesh.solid, gs, lam_s, mu_s, qmu, qka)
        for n in ("DsDeta_over_J", "DzDeta_over_J", "DsDxi_over_J", "DzDxi_over_J"):
            att_d[n + "_cg4"] = _f32(cg4(pw_s[n]))
        att_d["coarse_grained"] = att.coarse_grained
        att_d["do_corr_lowq"] = att.do_corr_lowq
        att_d["n_sls"] = att.n_sls

    solid = solid_stiffness_terms(src_type, mesh.solid, gs, lam_s, mu_s, xi_s, phi_s, eta_s,
                                  anel, basis)
    fluid = fluid_stiffness_terms(src_type, mesh.fluid, gf, rho_f, basis) if mesh.nel_fluid else {}
    bdry = sf_boundary_terms(mesh, gs) if mesh.nel_bdry else np.zeros((2, 0, 5), np.float32)

    # free-surface mask (time_evol_wave.F90:1615-1630)
    fsm = np.ones((mesh.nel_fluid, 5, 5), dtype=np.float32)
    if mesh.nel_fluid:
        rr = np.broadcast_to(gf.r[:, :, None], fsm.shape)
        fsm[rr > spec.router - 1.0] = 0.0

    nelsrc, ielsrc, st = compute_source_terms(mesh, source, pw_s)
    if time_scheme == "newmark2":
        stf = compute_stf(source, niter, deltat)
    else:
        stf = np.zeros(niter, dtype=np.float32)

    if rec_colat_deg is None:
        rec_colat_deg = np.linspace(5.0, 175.0, 18)
    rec = surface_receivers(mesh, rec_colat_deg)

    prob = Problem(
        mesh=mesh, src_type=src_type, src_order=SRC_ORDER[src_type], time_scheme=time_scheme,
        deltat=float(deltat), niter=int(niter), seis_it=int(seis_it), strain_it=int(strain_it),
        anel=anel, solid=solid, fluid=fluid, inv_mass_rho=_f32(inv_mass_rho),
        inv_mass_fluid=_f32(inv_mass_fluid), fluid_free_surface_mask=fsm,
        inv_rho_fluid=_f32(inv_rho_fluid), pw_solid=pw_s, pw_fluid=pw_f, bdry_matr=bdry,
        att=att_d, att_model=att if anel else None, source=source, nelsrc=nelsrc,
        ielsrc=ielsrc, source_term_el=st, stf=stf, recfile_el=rec["recfile_el"],
        rec_index=rec["index"])
    if dump:
        prob.kwf = kwf_maps(mesh)
    return prob


def kwf_maps(mesh: LocalMesh) -> Dict[str, np.ndarray]:
    """Wavefield-dump point set for `displ_only` (the whole mesh: get_mesh.f90:91-96),
    de-duplicated by global number with first-visit-wins in solid-then-fluid, element,
    jpol, ipol order (SOLVER/meshes_io.F90:489-640).  Returns kwf_mask (nel,5,5) int32
    0/1 and mapping_ijel_ikwf (nel,5,5) 1-based into [solid | fluid] points, for
    nel = nel_solid + nel_fluid."""
    ns, nf = mesh.nel_solid, mesh.nel_fluid
    out_mask = np.zeros((ns + nf, 5, 5), dtype=np.int32)
    out_map = np.zeros((ns + nf, 5, 5), dtype=np.int32)
    base = 0
    counts = []
    for dom, nel, ig in (("s", ns, mesh.igloc_solid), ("f", nf, mesh.igloc_fluid)):
        if nel == 0:
            counts.append(0)
            continue
        # visiting order of the reference: iel, jpol, ipol (= memory order)
        order = ig.reshape(-1).astype(np.int64)
        uniq, first = np.unique(order, return_index=True)
        rank_of_first = np.argsort(np.argsort(first))      # visit rank of each unique id
        lut = np.zeros(uniq.max() + 1, dtype=np.int64)
        lut[uniq] = rank_of_first + 1
        is_first = np.zeros(order.size, dtype=bool)
        is_first[first] = True
        off = 0 if dom == "s" else ns
        out_mask[off:off + nel] = is_first.reshape(nel, 5, 5)
        out_map[off:off + nel] = (lut[order] + base).reshape(nel, 5, 5)
        counts.append(uniq.size)
        base += uniq.size
    return {"kwf_mask": out_mask, "mapping_ijel_ikwf": out_map,
            "npoint_solid_kwf": counts[0], "npoint_fluid_kwf": counts[1]}
