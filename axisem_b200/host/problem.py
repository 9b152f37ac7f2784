"""Assemble one rank's complete set of time-loop inputs (what `prepare_waves`,
SOLVER/time_evol_wave.F90:47-225, leaves in the Fortran modules) for a synthetic mesh.

`build_problem` is the stand-in for the Fortran host in this repository: it produces
exactly the arrays the C ABI (`include/axisem_b200.h`) takes, in the reference's layouts
and index conventions.  The same `Problem` feeds the CUDA library, the CPU oracle and the
benchmark, so parity tests compare like with like.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
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
    unassem_mass_rho_solid: Optional[np.ndarray] = None   # dump_energy (def_precomp_terms.f90:745-751)
    unassem_mass_lam_fluid: Optional[np.ndarray] = None   # (:812-815)
    fluid_src: bool = False            # source inside the fluid (have_src in the fluid: add_source_fl)
    kwf: Optional[Dict[str, np.ndarray]] = None
    dump_type: str = "displ_only"      # data_io%dump_type: displ_only | strain_only | fullfields
    dump_block: tuple = (0, 4, 0, 4)   # ibeg, iend, jbeg, jend (fullfields; parameters.F90:400-403)
    xdmf: Optional[Dict] = None        # dump_xdmf: the maps of host/xdmf.py:xdmf_maps + "snap_it"

    @property
    def num_rec(self):
        return int(self.recfile_el.shape[0])


def _subset(es: ElementSet, idx) -> ElementSet:
    return ElementSet(es.it[idx], es.ir[idx], es.th_a[idx], es.th_b[idx], es.r_a[idx],
                      es.r_b[idx], es.axis[idx], es.north[idx], es.layer[idx])


def _assemble_columns(spec: MeshSpec, it0: int, it1: int, fluid: bool, own_val: np.ndarray,
                      ghost_val) -> np.ndarray:
    """Direct-stiffness-sum a per-point quantity over the *whole* mesh and return it on
    this rank's elements (def_mass_matrix_k calls pdistsum_* once, :756, :820).
    `own_val` is (ncol*nrd, 5, 5) in (it, ir) order; `ghost_val(c)` returns the values of
    the elements of column c (a neighbouring slice) for the cross-rank contributions."""
    irs = np.nonzero(spec.fluid_ir == fluid)[0]
    nrd = irs.size
    if nrd == 0 or own_val.shape[0] == 0:
        return own_val
    nrn = spec.nrnode_fluid if fluid else spec.nrnode_solid
    lo = max(it0 - 1, 0)
    hi = min(it1 + 1, spec.ntheta)
    dense = np.zeros(((hi - lo) * 4 + 1, nrn), dtype=np.float64)   # order-independent sums
    rb = spec.rbase[irs]
    half = spec.ntheta // 2

    def scatter(cols, val):          # cols: array of it, val: (ncols, nrd, 5, 5) float32
        for north in (True, False):
            m = (cols < half) == north
            if not m.any():
                continue
            c = cols[m] - lo
            v = val[m]
            for j in range(5):
                for i in range(5):
                    ip, jp = (i, j) if north else (4 - i, 4 - j)
                    dense[(4 * c + ip)[:, None], (rb + jp)[None, :]] += v[:, :, j, i]

    def gather(cols):
        out = np.empty((cols.size, nrd, 5, 5), dtype=np.float64)
        for north in (True, False):
            m = (cols < half) == north
            if not m.any():
                continue
            c = cols[m] - lo
            for j in range(5):
                for i in range(5):
                    ip, jp = (i, j) if north else (4 - i, 4 - j)
                    out[np.nonzero(m)[0][:, None], np.arange(nrd)[None, :], j, i] = \
                        dense[(4 * c + ip)[:, None], (rb + jp)[None, :]]
        return out

    cols = np.arange(it0, it1)
    scatter(cols, own_val.astype(np.float32).reshape(cols.size, nrd, 5, 5))
    for c in (it0 - 1, it1):
        if 0 <= c < spec.ntheta:
            scatter(np.array([c]), ghost_val(c).astype(np.float32).reshape(1, nrd, 5, 5))
    return gather(cols).reshape(-1, 5, 5)


def _element_arrays(spec: MeshSpec, basis: SpectralBasis, cols: np.ndarray, src_type: str,
                    anel: bool, att: Optional[AttenuationModel], deltat: float) -> Dict:
    """Every per-element array of the elements in theta columns `cols` (both domains)."""
    IT, IR = np.meshgrid(cols, np.arange(spec.nr), indexing="ij")
    IT = IT.reshape(-1)
    IR = IR.reshape(-1)
    fl = spec.fluid_ir[IR]
    es_s = make_elements(spec, IT[~fl], IR[~fl])
    es_f = make_elements(spec, IT[fl], IR[fl])
    out: Dict = {}
    gs = geometry(es_s, basis)
    rho_s, lam_s, mu_s, xi_s, phi_s, eta_s, _, qmu, qka = material(spec, es_s, gs)
    out["mass_s"] = _f32(rho_s * gs.massmat_k)
    out["pw_s"] = pointwise_derivative_terms(es_s, gs)
    if anel:
        att_d, lam_s, mu_s = attenuation_terms(att, deltat, es_s, gs, lam_s, mu_s, qmu, qka)
        for n in ("DsDeta_over_J", "DzDeta_over_J", "DsDxi_over_J", "DzDxi_over_J"):
            att_d[n + "_cg4"] = _f32(cg4(out["pw_s"][n]))
        out["att"] = att_d
    out["solid"] = solid_stiffness_terms(src_type, es_s, gs, lam_s, mu_s, xi_s, phi_s, eta_s,
                                         anel, basis)
    if es_f.nel:
        gf = geometry(es_f, basis)
        rho_f, lam_f, *_ = material(spec, es_f, gf)
        out["mass_f"] = _f32(gf.massmat_k / lam_f)
        out["inv_rho_fluid"] = _f32(1.0 / rho_f)
        out["pw_f"] = pointwise_derivative_terms(es_f, gf)
        out["fluid"] = fluid_stiffness_terms(src_type, es_f, gf, rho_f, basis)
        fsm = np.ones((es_f.nel, 5, 5), dtype=np.float32)
        rr = np.broadcast_to(gf.r[:, :, None], fsm.shape)
        fsm[rr > spec.router - 1.0] = 0.0                     # time_evol_wave.F90:1615-1630
        out["fsm"] = fsm
    return out


def _concat(parts, key):
    if isinstance(parts[0][key], dict):
        keys = parts[0][key].keys()
        res = {}
        for k in keys:
            v0 = parts[0][key][k]
            if isinstance(v0, np.ndarray) and v0.ndim >= 1 and v0.dtype == np.float32:
                res[k] = np.concatenate([p[key][k] for p in parts], axis=0)
            else:
                res[k] = v0          # rank-independent scalars / small vectors
        return res
    return np.concatenate([p[key] for p in parts], axis=0)


def build_problem(spec: MeshSpec, source: Optional[SourceParams] = None, *, rank: int = 0,
                  nranks: int = 1, anel: bool = False, att: Optional[AttenuationModel] = None,
                  time_scheme: str = "newmark2", niter: int = 100, seis_it: int = 1,
                  strain_it: int = 0, courant: float = 0.6, deltat: Optional[float] = None,
                  rec_colat_deg=None, dump: bool = False, energy: bool = False, chunk_cols: int = 32,
                  threads: Optional[int] = None, dump_type: str = "displ_only",
                  dump_block=(0, 4, 0, 4), nranks_r: int = 1, snap_it: int = 0,
                  xdmf_opts: Optional[Dict] = None) -> Problem:
    assert time_scheme in TIME_SCHEMES
    source = source or SourceParams()
    src_type = source.src_type1
    basis = SpectralBasis(spec.npol)
    mesh = build_rank(spec, rank, nranks, basis, nranks_r)
    if deltat is None:
        deltat = stable_timestep(spec, basis, courant)
        if time_scheme != "newmark2":
            deltat *= 1.5                      # time_evol_wave.F90:511
    if anel:
        att = att or AttenuationModel()

    # ---- per-element arrays, built in chunks of theta columns on a thread pool ---------
    cols = np.arange(mesh.it0, mesh.it1)
    chunks = [cols[k:k + chunk_cols] for k in range(0, cols.size, chunk_cols)]
    work = lambda c: _element_arrays(spec, basis, c, src_type, anel, att, deltat)
    nthr = threads if threads is not None else min(len(chunks), os.cpu_count() or 1)
    if nthr > 1:
        with ThreadPoolExecutor(max_workers=nthr) as ex:
            parts = list(ex.map(work, chunks))
    else:
        parts = [work(c) for c in chunks]
    has_fluid = mesh.nel_fluid > 0
    solid = _concat(parts, "solid")
    pw_s = _concat(parts, "pw_s")
    att_d = None
    if anel:
        att_d = _concat(parts, "att")
        att_d["coarse_grained"] = att.coarse_grained
        att_d["do_corr_lowq"] = att.do_corr_lowq
        att_d["n_sls"] = att.n_sls
    if has_fluid:
        fluid = _concat(parts, "fluid")
        pw_f = _concat(parts, "pw_f")
        inv_rho_fluid = _concat(parts, "inv_rho_fluid")
        fsm = _concat(parts, "fsm")
    else:
        fluid, pw_f = {}, {k: np.zeros((0, 5, 5), np.float32) for k in pw_s}
        inv_rho_fluid = np.zeros((0, 5, 5), np.float32)
        fsm = np.zeros((0, 5, 5), np.float32)

    # ---- mass matrices (before attenuation changes the moduli, as in the reference) ---
    def ghost(fluidflag):
        def f(c):
            irs = np.nonzero(spec.fluid_ir == fluidflag)[0]
            es = make_elements(spec, np.full(irs.size, c), irs)
            g = geometry(es, basis)
            m = material(spec, es, g)
            return (g.massmat_k / m[1]) if fluidflag else (m[0] * g.massmat_k)
        return f

    um_s = um_f = None
    if energy:
        um_s = _f32(_concat(parts, "mass_s") * (2.0 if src_type == "dipole" else 1.0))
        um_f = _f32(_concat(parts, "mass_f")) if has_fluid else np.zeros((0, 5, 5), np.float32)
    m_s = _assemble_columns(spec, mesh.it0, mesh.it1, False, _concat(parts, "mass_s"), ghost(False))
    inv_mass_rho = (1.0 / m_s.astype(np.float64))
    if src_type == "dipole":
        inv_mass_rho = 0.5 * inv_mass_rho                 # def_precomp_terms.f90:773
    if has_fluid:
        m_f = _assemble_columns(spec, mesh.it0, mesh.it1, True, _concat(parts, "mass_f"), ghost(True))
        inv_mass_fluid = 1.0 / m_f.astype(np.float64)
    else:
        inv_mass_fluid = np.zeros((0, 5, 5))
    del parts
    if nranks_r > 1:
        # the arrays above cover whole columns (the mass matrix is assembled over them); keep this
        # rank's radial block
        IT, IR = np.meshgrid(cols, np.arange(spec.nr), indexing="ij")
        IR = IR.reshape(-1)
        fl = spec.fluid_ir[IR]
        sel_s = (IR[~fl] >= mesh.ir0) & (IR[~fl] < mesh.ir1)
        sel_f = (IR[fl] >= mesh.ir0) & (IR[fl] < mesh.ir1)
        n_s, n_f = sel_s.size, sel_f.size

        def cut(a, sel, n):
            return a[sel] if isinstance(a, np.ndarray) and a.ndim >= 1 and a.shape[0] == n else a

        solid = {k: cut(v, sel_s, n_s) for k, v in solid.items()}
        pw_s = {k: cut(v, sel_s, n_s) for k, v in pw_s.items()}
        if att_d is not None:
            att_d = {k: cut(v, sel_s, n_s) for k, v in att_d.items()}
        inv_mass_rho = inv_mass_rho[sel_s]
        if um_s is not None:
            um_s = um_s[sel_s]
        if has_fluid:
            fluid = {k: cut(v, sel_f, n_f) for k, v in fluid.items()}
            pw_f = {k: cut(v, sel_f, n_f) for k, v in pw_f.items()}
            inv_rho_fluid, fsm, inv_mass_fluid = inv_rho_fluid[sel_f], fsm[sel_f], inv_mass_fluid[sel_f]
            if um_f is not None:
                um_f = um_f[sel_f]
        if mesh.nel_fluid == 0:
            fluid, pw_f = {}, {k: np.zeros((0, 5, 5), np.float32) for k in pw_s}
            inv_rho_fluid = fsm = np.zeros((0, 5, 5), np.float32)
            inv_mass_fluid = np.zeros((0, 5, 5))
        assert inv_mass_rho.shape[0] == mesh.nel_solid

    if mesh.nel_bdry:
        bidx = mesh.bdry_solid_el - 1
        sub = _subset(mesh.solid, bidx)
        bdry = sf_boundary_terms(mesh, geometry(sub, basis), sub)
    else:
        bdry = np.zeros((2, 0, 5), np.float32)

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
        inv_rho_fluid=inv_rho_fluid, pw_solid=pw_s, pw_fluid=pw_f, bdry_matr=bdry,
        att=att_d, att_model=att if anel else None, source=source, nelsrc=nelsrc,
        ielsrc=ielsrc, source_term_el=st, stf=stf, recfile_el=rec["recfile_el"],
        rec_index=rec["index"])
    if dump:
        prob.kwf = kwf_maps(mesh)
        prob.dump_type = dump_type
        prob.dump_block = tuple(int(v) for v in dump_block)
    if snap_it > 0:
        # SAVE_SNAPSHOTS with SNAPSHOTS_FORMAT xdmf: snap_it = floor(snap_dt / deltat) (parameters.F90:944)
        from .xdmf import xdmf_maps
        prob.xdmf = xdmf_maps(mesh, **(xdmf_opts or {}))
        prob.xdmf["snap_it"] = int(snap_it)
    prob.unassem_mass_rho_solid, prob.unassem_mass_lam_fluid = um_s, um_f
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
