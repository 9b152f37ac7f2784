"""Synthetic PREM-type 2-D (s,z) half-disc meshes in the reference's mesh-database
conventions, with a theta-slice domain decomposition.

The reference's MESHER (16.7 kLoC Fortran, out of scope) writes one `meshdb.datNNNN` per
rank (MESHER/pdb.f90:2191-2389) which the SOLVER ingests (SOLVER/get_mesh.f90:47-409).
This module produces the *same arrays* (names follow SOLVER/data_mesh.f90 and
data_comm.f90) for a structured spherical-shell mesh:

  * all elements are concentric "curved" spheroidal elements
    (SOLVER/analytic_spheroid_mapping.f90:40-107): theta linear in xi, r linear in eta;
  * southern-hemisphere elements are mirrored (xi=-1 on the larger colatitude, eta=-1
    on the larger radius) so that the Jacobian stays positive and axial elements always
    have ipol=0 on the axis, as in the reference;
  * axial elements use GLJ(0,1) nodes in xi (SOLVER/def_precomp_terms.f90:201-223);
  * global numbers are per domain (solid / fluid) and per rank ("glocal"), element-local
    linear index ipt = (iel-1)*25 + jpol*5 + ipol + 1 (SOLVER/commun.F90:303);
  * rank r owns a contiguous block of theta columns; neighbours exchange partial sums
    on the shared column of GLL points (SOLVER/data_comm.f90:36-71).

Deviations from a real AxiSEM mesh (documented in DESIGN.md): no inner cube (the sphere
is hollow below r_min with a free inner surface), no lateral coarsening layers.
All integer maps are 1-based int32 exactly as the Fortran host would pass them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .model import Layer, prem_layers, evaluate_layer
from .spectral import SpectralBasis

NP1 = 5          # npol + 1
NPT = 25         # points per element


@dataclass
class MeshSpec:
    """Global (all-rank) description of a structured shell mesh."""
    ntheta: int                      # lateral elements over [0, pi]; even
    layers: List[Layer]
    nrad: List[int]                  # radial elements per layer
    npol: int = 4

    def __post_init__(self):
        assert self.npol == 4, "hot path is npol=4 (SURVEY.md section 8)"
        assert self.ntheta % 2 == 0 and self.ntheta >= 2
        assert len(self.layers) == len(self.nrad)
        edges = [self.layers[0].r_bot]
        lay = []
        for k, (L, n) in enumerate(zip(self.layers, self.nrad)):
            e = np.linspace(L.r_bot, L.r_top, n + 1)
            edges.extend(e[1:].tolist())
            lay.extend([k] * n)
        self.r_edges = np.array(edges, dtype=np.float64)
        self.layer_of_ir = np.array(lay, dtype=np.int64)
        self.nr = len(lay)
        self.fluid_ir = np.array([self.layers[k].fluid for k in lay], dtype=bool)
        self.theta_edges = np.linspace(0.0, np.pi, self.ntheta + 1)
        self.theta_edges[-1] = np.pi
        self.router = self.layers[-1].r_top
        # per-domain radial numbering: rbase[ir] + j', runs of equal domain share nodes
        self.rbase = np.zeros(self.nr, dtype=np.int64)
        nsol = nflu = 0
        for ir in range(self.nr):
            new_run = ir == 0 or self.fluid_ir[ir] != self.fluid_ir[ir - 1]
            if self.fluid_ir[ir]:
                if new_run and nflu > 0:
                    nflu += 1
                self.rbase[ir] = nflu
                nflu += 4
            else:
                if new_run and nsol > 0:
                    nsol += 1
                self.rbase[ir] = nsol
                nsol += 4
        self.nrnode_solid = nsol + 1 if (~self.fluid_ir).any() else 0
        self.nrnode_fluid = nflu + 1 if self.fluid_ir.any() else 0

    @property
    def nelem(self):
        return self.ntheta * self.nr


def prem_mesh_spec(ntheta: int, nr_target: int, anisotropic: bool = False,
                   r_min_km: float = 400.0) -> MeshSpec:
    """PREM-type layering with about `nr_target` radial elements, distributed over the
    layers in proportion to thickness / (local S or P wavelength)."""
    layers = prem_layers(anisotropic=anisotropic, r_min_km=r_min_km)
    cost = []
    for L in layers:
        rm = 0.5 * (L.r_bot + L.r_top)
        rho, lam, mu, *_ = evaluate_layer(L, np.array([rm]))
        v = np.sqrt(mu[0] / rho[0]) if not L.fluid else np.sqrt(lam[0] / rho[0])
        cost.append((L.r_top - L.r_bot) / v)
    cost = np.array(cost)
    nrad = np.maximum(1, np.round(cost / cost.sum() * nr_target).astype(int)).tolist()
    return MeshSpec(ntheta=ntheta, layers=layers, nrad=nrad)


@dataclass
class ElementSet:
    """Geometry of a list of elements of one domain (solid or fluid)."""
    it: np.ndarray          # global theta column
    ir: np.ndarray          # global radial index
    th_a: np.ndarray        # colatitude at xi=-1
    th_b: np.ndarray        # colatitude at xi=+1
    r_a: np.ndarray         # radius at eta=-1
    r_b: np.ndarray         # radius at eta=+1
    axis: np.ndarray        # bool
    north: np.ndarray       # bool
    layer: np.ndarray       # index into spec.layers

    @property
    def nel(self):
        return int(self.it.size)


def make_elements(spec: MeshSpec, it: np.ndarray, ir: np.ndarray) -> ElementSet:
    it = np.asarray(it, dtype=np.int64)
    ir = np.asarray(ir, dtype=np.int64)
    north = it < spec.ntheta // 2
    t0 = spec.theta_edges[it]
    t1 = spec.theta_edges[it + 1]
    r0 = spec.r_edges[ir]
    r1 = spec.r_edges[ir + 1]
    th_a = np.where(north, t0, t1)
    th_b = np.where(north, t1, t0)
    r_a = np.where(north, r0, r1)
    r_b = np.where(north, r1, r0)
    axis = (it == 0) | (it == spec.ntheta - 1)
    return ElementSet(it, ir, th_a, th_b, r_a, r_b, axis, north, spec.layer_of_ir[ir])


@dataclass
class HaloSide:
    """One domain's (solid or fluid) halo description, reference naming
    (SOLVER/data_comm.f90:36-71).  Send and receive lists are identical
    (SOLVER/get_mesh.f90:303-310)."""
    nmsg: int = 0
    list_peer: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    sizemsg: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    glocal_index_msg: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), np.int32))
    num_comm_gll: int = 0
    glob2el: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.int32))


@dataclass
class LocalMesh:
    """Everything a rank's SOLVER holds about its piece of the mesh (data_mesh.f90)."""
    spec: MeshSpec
    basis: SpectralBasis
    rank: int
    nranks: int
    it0: int
    it1: int
    nranks_r: int               # radial blocks (rank = theta_block * nranks_r + radial_block)
    ir0: int
    ir1: int
    solid: ElementSet
    fluid: ElementSet
    nel_solid: int
    nel_fluid: int
    igloc_solid: np.ndarray
    igloc_fluid: np.ndarray
    nglob_solid: int
    nglob_fluid: int
    axis_solid: np.ndarray      # int32 0/1 (Fortran logical(4))
    axis_fluid: np.ndarray
    ax_el_solid: np.ndarray     # 1-based
    ax_el_fluid: np.ndarray
    nel_bdry: int
    bdry_solid_el: np.ndarray   # 1-based
    bdry_fluid_el: np.ndarray
    bdry_jpol_solid: np.ndarray  # 0-based pol index, as in the reference
    bdry_jpol_fluid: np.ndarray
    bdry_above: np.ndarray      # bool: solid above fluid (sign of bdry_matr)
    halo_solid: HaloSide
    halo_fluid: HaloSide
    # global ids (for cross-rank tests): unique over the whole mesh, per domain
    gid_solid: np.ndarray
    gid_fluid: np.ndarray

    # ---- coordinates ---------------------------------------------------------------
    def coords(self, dom: str):
        """theta[e,i], r[e,j], s[e,j,i], z[e,j,i] at the element's collocation points
        (GLJ in xi for axial elements)."""
        es = self.solid if dom == "solid" else self.fluid
        return element_coords(es, self.basis)


def element_coords(es: ElementSet, basis: SpectralBasis):
    xi = np.where(es.axis[:, None], basis.xi_k[None, :], basis.eta[None, :])   # (nel,5)
    th = 0.5 * ((1.0 - xi) * es.th_a[:, None] + (1.0 + xi) * es.th_b[:, None])
    eta = basis.eta[None, :]
    r = 0.5 * ((1.0 - eta) * es.r_a[:, None] + (1.0 + eta) * es.r_b[:, None])
    sin_t = np.sin(th)
    cos_t = np.cos(th)
    # exact zeros on the axis (the reference masks |s| < min_distance_dim,
    # analytic_spheroid_mapping.f90:62)
    on_axis = es.axis
    sin_t[on_axis, 0] = 0.0
    cos_t[on_axis, 0] = np.sign(cos_t[on_axis, 0])
    s = r[:, :, None] * sin_t[:, None, :]
    z = r[:, :, None] * cos_t[:, None, :]
    return xi, th, r, s, z, sin_t, cos_t


def _domain_numbering(spec: MeshSpec, es: ElementSet, it0: int, fluid: bool):
    """igloc (1-based, rank-local), the rank's radial node range (first node, count) and a
    mesh-global id for every element-local point.  The local numbers run over the radial nodes
    the rank's elements of this domain touch (the whole column for theta-only slices)."""
    nrn = spec.nrnode_fluid if fluid else spec.nrnode_solid
    nel = es.nel
    if nel == 0:
        return np.zeros(0, np.int32), (0, 0), np.zeros(0, np.int64)
    i = np.arange(NP1)
    ip = np.where(es.north[:, None], i[None, :], 4 - i[None, :])          # (nel,5) theta
    jp = ip                                                                 # same flip in r
    tnode = 4 * es.it[:, None] + ip                                         # global
    rnode = spec.rbase[es.ir][:, None] + jp
    gid = tnode[:, None, :] * nrn + rnode[:, :, None]                       # (nel,j,i)
    r0 = int(spec.rbase[es.ir].min())
    nloc = int(spec.rbase[es.ir].max()) + 4 - r0 + 1
    loc = (tnode - 4 * it0)[:, None, :] * nloc + (rnode[:, :, None] - r0) + 1
    return loc.reshape(-1).astype(np.int32), (r0, nloc), gid.reshape(-1)


def radial_blocks(spec: MeshSpec, nranks_r: int):
    """ir ranges of a radial decomposition into nranks_r blocks of about equal element counts
    (MESHER/parallelization.f90:157-165 cuts in radius as well); a cut never falls on a solid/fluid
    boundary, so that every S/F boundary pair stays inside one rank."""
    cuts = [0]
    for k in range(1, nranks_r):
        c = int(round(k * spec.nr / nranks_r))
        while 0 < c < spec.nr and spec.fluid_ir[c - 1] != spec.fluid_ir[c]:
            c += 1
        assert cuts[-1] < c < spec.nr, "too many radial blocks for this mesh"
        cuts.append(c)
    cuts.append(spec.nr)
    return [(cuts[k], cuts[k + 1]) for k in range(nranks_r)]


def build_rank(spec: MeshSpec, rank: int = 0, nranks: int = 1,
               basis: Optional[SpectralBasis] = None, nranks_r: int = 1) -> LocalMesh:
    """The piece of the mesh owned by `rank` of a theta (x radius) decomposition
    (MESHER/parallelization.f90:68-112 gives equal element counts per slice, :157-165 the radial
    cuts).  rank = theta_block * nranks_r + radial_block; up to eight neighbours per rank."""
    basis = basis or SpectralBasis(spec.npol)
    assert nranks % nranks_r == 0
    nth = nranks // nranks_r
    assert spec.ntheta % nth == 0, "ntheta must be divisible by the number of theta slices"
    bt, br = divmod(rank, nranks_r)
    ncol = spec.ntheta // nth
    it0, it1 = bt * ncol, (bt + 1) * ncol
    blocks = radial_blocks(spec, nranks_r)
    ir0, ir1 = blocks[br]
    cols = np.arange(it0, it1)
    IT, IR = np.meshgrid(cols, np.arange(ir0, ir1), indexing="ij")         # ir fastest
    IT = IT.reshape(-1)
    IR = IR.reshape(-1)
    fl = spec.fluid_ir[IR]
    solid = make_elements(spec, IT[~fl], IR[~fl])
    fluid = make_elements(spec, IT[fl], IR[fl])

    ig_s, (r0_s, nloc_s), gid_s = _domain_numbering(spec, solid, it0, False)
    ig_f, (r0_f, nloc_f), gid_f = _domain_numbering(spec, fluid, it0, True)
    nglob_s = (4 * ncol + 1) * nloc_s
    nglob_f = (4 * ncol + 1) * nloc_f

    # local element index by (it, ir)
    idx_s = -np.ones((ncol, spec.nr), dtype=np.int64)
    idx_s[solid.it - it0, solid.ir] = np.arange(solid.nel)
    idx_f = -np.ones((ncol, spec.nr), dtype=np.int64)
    idx_f[fluid.it - it0, fluid.ir] = np.arange(fluid.nel)

    # ---- solid/fluid boundary pairs (data_mesh.f90:106-110) ------------------------
    bs, bf, js, jf, above = [], [], [], [], []
    for ir in range(ir0, ir1 - 1):
        lo_f, hi_f = spec.fluid_ir[ir], spec.fluid_ir[ir + 1]
        if lo_f == hi_f:
            continue
        for c in range(ncol):
            north = (it0 + c) < spec.ntheta // 2
            if hi_f:      # fluid above solid (e.g. ICB)
                es, ef = idx_s[c, ir], idx_f[c, ir + 1]
                j_s = 4 if north else 0     # top edge of the solid element
                j_f = 0 if north else 4     # bottom edge of the fluid element
                ab = False
            else:         # solid above fluid (e.g. CMB)
                es, ef = idx_s[c, ir + 1], idx_f[c, ir]
                j_s = 0 if north else 4
                j_f = 4 if north else 0
                ab = True
            bs.append(es + 1)
            bf.append(ef + 1)
            js.append(j_s)
            jf.append(j_f)
            above.append(ab)
    nel_bdry = len(bs)

    # ---- halo ---------------------------------------------------------------------
    def halo(es: ElementSet, idx, fluid, r0, nloc, igloc):
        """Messages to the (up to eight) neighbouring blocks: the shared column of radial nodes
        for theta neighbours, the shared row of theta nodes for radial neighbours (only where the
        elements across the cut belong to the same domain, i.e. share nodes), the single corner
        node for diagonal ones.  Lists are ordered by node index, the same on both sides."""
        h = HaloSide()
        if es.nel == 0 or nranks == 1:
            return h
        dom = spec.fluid_ir == fluid
        # is the node row at the bottom / top of this block shared with the block below / above?
        row_shared = {-1: br > 0 and dom[ir0 - 1] and dom[ir0], 1: br < nranks_r - 1 and dom[ir1 - 1] and dom[ir1]}
        rloc_row = {-1: int(spec.rbase[ir0]) - r0 if dom[ir0] else -1,
                    1: int(spec.rbase[ir1 - 1]) + 4 - r0 if dom[ir1 - 1] else -1}
        tloc_col = {-1: 0, 1: 4 * ncol}
        peers, lists = [], []
        for dt in (-1, 0, 1):
            for dr in (-1, 0, 1):
                if (dt, dr) == (0, 0) or not (0 <= bt + dt < nth) or not (0 <= br + dr < nranks_r):
                    continue
                if dr != 0 and not row_shared[dr]:
                    continue
                peer = (bt + dt) * nranks_r + (br + dr)
                if dr == 0:
                    # the peer has the same radial block: the whole column of this domain's nodes
                    ids = tloc_col[dt] * nloc + np.arange(nloc) + 1
                elif dt == 0:
                    ids = np.arange(4 * ncol + 1) * nloc + rloc_row[dr] + 1
                else:
                    ids = np.array([tloc_col[dt] * nloc + rloc_row[dr] + 1])
                # nodes of the column that no element of this rank touches (another domain's gap) are not sent
                ids = ids[np.isin(ids, igloc)]
                if ids.size:
                    peers.append(peer)
                    lists.append(ids)
        h.nmsg = len(peers)
        if h.nmsg == 0:
            return h
        h.list_peer = np.array(peers, dtype=np.int32)
        h.sizemsg = np.array([len(l) for l in lists], dtype=np.int32)
        m = max(len(l) for l in lists)
        h.glocal_index_msg = np.zeros((h.nmsg, m), dtype=np.int32)
        for k, l in enumerate(lists):
            h.glocal_index_msg[k, :len(l)] = l
        # glob2el (def_grid.f90:95-180): element-local points, in memory order, whose number is sent
        sent = np.unique(np.concatenate(lists))
        hit = np.nonzero(np.isin(igloc, sent))[0]
        e, q = hit // NPT, hit % NPT
        h.glob2el = np.stack([q % NP1, q // NP1, e + 1], axis=1).astype(np.int32)
        h.num_comm_gll = h.glob2el.shape[0]
        return h

    halo_s = halo(solid, idx_s, False, r0_s, nloc_s, ig_s)
    halo_f = halo(fluid, idx_f, True, r0_f, nloc_f, ig_f)

    return LocalMesh(
        spec=spec, basis=basis, rank=rank, nranks=nranks, it0=it0, it1=it1, nranks_r=nranks_r, ir0=ir0, ir1=ir1,
        solid=solid, fluid=fluid, nel_solid=solid.nel, nel_fluid=fluid.nel,
        igloc_solid=ig_s, igloc_fluid=ig_f, nglob_solid=nglob_s, nglob_fluid=nglob_f,
        axis_solid=solid.axis.astype(np.int32), axis_fluid=fluid.axis.astype(np.int32),
        ax_el_solid=(np.nonzero(solid.axis)[0] + 1).astype(np.int32),
        ax_el_fluid=(np.nonzero(fluid.axis)[0] + 1).astype(np.int32),
        nel_bdry=nel_bdry,
        bdry_solid_el=np.array(bs, dtype=np.int32), bdry_fluid_el=np.array(bf, dtype=np.int32),
        bdry_jpol_solid=np.array(js, dtype=np.int32), bdry_jpol_fluid=np.array(jf, dtype=np.int32),
        bdry_above=np.array(above, dtype=bool),
        halo_solid=halo_s, halo_fluid=halo_f, gid_solid=gid_s, gid_fluid=gid_f)


def surface_receivers(mesh: LocalMesh, colat_deg) -> Dict[str, np.ndarray]:
    """Nearest surface GLL point to each requested colatitude, kept only if this rank
    owns it (SOLVER/seismograms.f90:235-639 does the same search; ties go to the lower
    rank).  Returns recfile_el(num_rec,3) = (iel 1-based solid-local, ipol, jpol) as in
    data_mesh.f90:138 plus the indices of the stations kept."""
    spec = mesh.spec
    es = mesh.solid
    top = np.nonzero(es.ir == spec.nr - 1)[0]
    xi, th, *_ = element_coords(es, mesh.basis)
    rec, keep = [], []
    if top.size == 0:                # a radial block below the surface holds no receivers
        return {"recfile_el": np.zeros((0, 3), dtype=np.int32), "index": np.zeros(0, dtype=np.int64)}
    colat = np.deg2rad(np.atleast_1d(np.asarray(colat_deg, dtype=np.float64)))
    lo = spec.theta_edges[mesh.it0]
    hi = spec.theta_edges[mesh.it1]
    for k, c in enumerate(colat):
        owner_ok = (lo <= c < hi) or (mesh.it1 == spec.ntheta and c == hi)
        if not owner_ok:
            continue
        d = np.abs(th[top, :] - c)
        e_loc, i = np.unravel_index(np.argmin(d), d.shape)
        e = top[e_loc]
        j = 4 if es.north[e] else 0
        rec.append((e + 1, i, j))
        keep.append(k)
    return {"recfile_el": np.array(rec, dtype=np.int32).reshape(-1, 3),
            "index": np.array(keep, dtype=np.int64)}
