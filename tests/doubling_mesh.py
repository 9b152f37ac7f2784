"""A conforming mesh with a lateral coarsening ("doubling") layer, written as a MESHER database — test
fixture for the native reader and pre-computation (axisem_b200/hostcxx/meshdb.cpp, mapping.cpp,
precomp.cpp) on a mesh that is not a theta x r grid.

Shell r_min .. router, all solid (the mantle of prem_iso_light above a free inner surface at the CMB): `nth` columns above the layer, `nth`/2
below it, and between them the 4-to-2 template the mesher's coarsening layers are made of — six elements
per period with one circular and one straight side (eltype semino / semiso of analytic_semi_mapping.f90),
straight diagonals in between:

    y=1   T0----T1----T2----T3----T4        q1 [B0 P  T1 T0] semino   q4 [C  Q  T3 T2] semino
          | q1  |  q3 |  q4 | q6  |         q2 [B0 B2 C  P ] semiso   q5 [B2 B4 Q  C ] semiso
    y=.5  |     P-----C-----Q     |         q3 [P  C  T2 T1] semino   q6 [Q  B4 T4 T3] semino
          |   /   q2  |  q5   \\   |
    y=0   B0----------B2----------B4        (corners counter-clockwise from (xi, eta) = (-1, -1))

The axis cuts the tiling along B2-C-T2, so that axial elements have their xi = -1 edge on it (GLJ points in
xi on both sides of every shared xi-edge).  Southern elements are turned by 180 degrees (xi = -1 towards
the south axis, eta = -1 on the larger radius) as in the mesher; that swaps semino and semiso.  Global
numbers are topological: corners by node, edge points by the edge's end nodes, so nothing depends on
coordinates agreeing to the last bit."""
import struct

import numpy as np

TEMPLATE = [
    ([(0, 0), (1, .5), (1, 1), (0, 1)], "semino"),
    ([(0, 0), (2, 0), (2, .5), (1, .5)], "semiso"),
    ([(1, .5), (2, .5), (2, 1), (1, 1)], "semino"),
    ([(2, .5), (3, .5), (3, 1), (2, 1)], "semino"),
    ([(2, 0), (4, 0), (3, .5), (2, .5)], "semiso"),
    ([(3, .5), (4, 0), (4, 1), (3, 1)], "semino"),
]


def _sz(t, r):
    s = 0.0 if (t == 0.0 or t == np.pi) else r * np.sin(t)
    return (s, r * np.cos(t))


def build(nth=16, r_coarse=(3480e3, 3630e3, 4115e3, 4600e3), r_dbl=(4600e3, 4900e3),
          r_fine=(4900e3, 5250e3, 5600e3, 5701e3, 5771e3, 5971e3, 6151e3, 6291e3, 6371e3), doubling=True,
          cube_halfwidth=None):
    """-> dict with everything write_database() needs.  doubling=False: the same radial layering with `nth`
    columns everywhere (the comparison mesh).  cube_halfwidth = a: the sphere below r_coarse[0] is filled as
    the mesher fills it — a half square [0, a] x [-a, a] of `linear` elements and, around it, one ring of
    elements with a straight side on the square and a circular one on r_coarse[0] (semino / semiso)."""
    assert nth % 8 == 0
    th_f = np.linspace(0.0, np.pi, nth + 1)
    th_f[-1] = np.pi
    els = []                                        # (corners [(s, z)] x 4 counter-clockwise, eltype, coarsing, turn in the south)

    def regular(thetas, radii):
        for r0, r1 in zip(radii[:-1], radii[1:]):
            for t0, t1 in zip(thetas[:-1], thetas[1:]):
                els.append(([_sz(t0, r0), _sz(t1, r0), _sz(t1, r1), _sz(t0, r1)], "curved", False, True))

    th_c = th_f[::2] if doubling else th_f
    if cube_halfwidth:
        a, nc = float(cube_halfwidth), (len(th_c) - 1) // 4
        assert 4 * nc == len(th_c) - 1 and a * np.sqrt(2.0) < r_coarse[0]
        h = a / nc
        for jz in range(2 * nc):
            for js in range(nc):
                s0, s1, z0, z1 = js * h, (js + 1) * h, -a + jz * h, -a + (jz + 1) * h
                els.append(([(s0, z0), (s1, z0), (s1, z1), (s0, z1)], "linear", False, False))
        # the square's boundary from the north axis to the south axis, one node per column edge of the shell
        bnd = [(k * h, a) for k in range(nc)] + [(a, a - k * h) for k in range(2 * nc)] + [(a - k * h, -a) for k in range(nc + 1)]
        for k in range(4 * nc):
            els.append(([bnd[k], bnd[k + 1], _sz(th_c[k + 1], r_coarse[0]), _sz(th_c[k], r_coarse[0])], "semino", False, True))
    if doubling:
        regular(th_c, r_coarse)
        r0, r1 = r_dbl
        for m in range(nth // 4 + 1):               # template x in [0, 4] on fine columns [4m - 2, 4m + 2]
            for corners, kind in TEMPLATE:
                cols = [4 * m - 2 + x for x, _ in corners]
                if min(cols) < 0 or max(cols) > nth:
                    continue
                els.append(([_sz(th_f[int(c)], r0 + y * (r1 - r0)) for c, (_, y) in zip(cols, corners)], kind, True, True))
        regular(th_f, r_fine)
    else:
        regular(th_f, tuple(r_coarse) + tuple(r_fine))
    return _finish(els, router=float(r_fine[-1]), rmin=0.0 if cube_halfwidth else float(r_coarse[0]))


def build_rows(rows, ncol_bottom, *, cube_halfwidth=None, fluid=None):
    """The general form: `rows` from the bottom up, (r0, r1, 'R') a regular row, (r0, r1, 'D') a coarsening
    row above which the number of columns is twice that below; `ncol_bottom` columns in the lowest row
    (four per side element of the inner square when cube_halfwidth is given, which fills the sphere below
    rows[0][0]).  fluid(r_mid) -> bool marks the fluid rows (regular rows only)."""
    els = []
    ncol = ncol_bottom
    th = np.linspace(0.0, np.pi, ncol + 1)
    th[-1] = np.pi
    if cube_halfwidth:
        a, nc = float(cube_halfwidth), ncol // 4
        assert 4 * nc == ncol and a * np.sqrt(2.0) < rows[0][0]
        h = a / nc
        for jz in range(2 * nc):
            for js in range(nc):
                s0, s1, z0, z1 = js * h, (js + 1) * h, -a + jz * h, -a + (jz + 1) * h
                els.append(([(s0, z0), (s1, z0), (s1, z1), (s0, z1)], "linear", False, False))
        bnd = [(k * h, a) for k in range(nc)] + [(a, a - k * h) for k in range(2 * nc)] + [(a - k * h, -a) for k in range(nc + 1)]
        for k in range(4 * nc):
            els.append(([bnd[k], bnd[k + 1], _sz(th[k + 1], rows[0][0]), _sz(th[k], rows[0][0])], "semino", False, True))
    for r0, r1, kind in rows:
        if kind == "R":
            for t0, t1 in zip(th[:-1], th[1:]):
                els.append(([_sz(t0, r0), _sz(t1, r0), _sz(t1, r1), _sz(t0, r1)], "curved", False, True))
        else:
            assert ncol % 4 == 0
            nth = 2 * ncol
            th_f = np.linspace(0.0, np.pi, nth + 1)
            th_f[-1] = np.pi
            for m in range(nth // 4 + 1):
                for corners, k2 in TEMPLATE:
                    cols = [4 * m - 2 + x for x, _ in corners]
                    if min(cols) < 0 or max(cols) > nth:
                        continue
                    els.append(([_sz(th_f[int(c)], r0 + y * (r1 - r0)) for c, (_, y) in zip(cols, corners)], k2, True, True))
            ncol, th = nth, th_f
    return _finish(els, router=float(rows[-1][1]), rmin=0.0 if cube_halfwidth else float(rows[0][0]), fluid=fluid)


def _finish(els, router, rmin, fluid=None):
    nelem = len(els)
    # orientation: southern elements of the shell are turned by 180 degrees
    corners, eltype, coarsing = [], [], []
    for c, kind, co, turn in els:
        if turn and np.mean([z for _, z in c]) < 0.0:
            c = [c[2], c[3], c[0], c[1]]
            kind = {"semino": "semiso", "semiso": "semino"}.get(kind, kind)
        corners.append(c)
        eltype.append(kind)
        coarsing.append(co)
    # nodes
    key = lambda s_, z_: (int(round(s_ * 64)), int(round(z_ * 64)))
    node_id = {}
    for c in corners:
        for s_, z_ in c:
            node_id.setdefault(key(s_, z_), len(node_id))
    cn = np.array([[node_id[key(s_, z_)] for s_, z_ in c] for c in corners])       # (nelem, 4)
    # global numbers of the 25 points of each element
    ig = np.zeros((nelem, 5, 5), dtype=np.int64)                                # [e, j, i]
    nxt = [len(node_id)]
    edge_pts = {}

    def edge(a, b):
        """the three interior points of edge a -> b, in that direction"""
        k = (min(a, b), max(a, b))
        if k not in edge_pts:
            edge_pts[k] = [nxt[0], nxt[0] + 1, nxt[0] + 2]
            nxt[0] += 3
        p = edge_pts[k]
        return p if a < b else p[::-1]

    for e in range(nelem):
        c0, c1, c2, c3 = cn[e]
        ig[e, 0, 0], ig[e, 0, 4], ig[e, 4, 4], ig[e, 4, 0] = c0, c1, c2, c3
        ig[e, 0, 1:4] = edge(c0, c1)
        ig[e, 4, 1:4] = edge(c3, c2)
        ig[e, 1:4, 0] = edge(c0, c3)
        ig[e, 1:4, 4] = edge(c1, c2)
        ig[e, 1:4, 1:4] = np.arange(nxt[0], nxt[0] + 9).reshape(3, 3)
        nxt[0] += 9
    # compress to 1..nglob in order of first appearance
    flat = ig.reshape(-1)
    _, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    order = np.argsort(np.argsort(first))
    igloc = (order[inv] + 1).astype(np.int32)
    nglob = int(igloc.max())
    # control nodes: corners and edge mid-points (on the circle for the circular sides), not shared between elements
    crd = np.zeros((8 * nelem, 2))
    for e, c in enumerate(corners):
        for k in range(4):
            pa, pb = np.array(c[k]), np.array(c[(k + 1) % 4])
            ra, rb = np.hypot(*pa), np.hypot(*pb)
            crd[8 * e + 2 * k] = pa
            mid = 0.5 * (pa + pb)
            circ = abs(ra - rb) < 1e-3 and ra > 0 and (eltype[e] == "curved" or (eltype[e] == "semino" and k == 2)
                                                       or (eltype[e] == "semiso" and k == 0))
            if circ and np.hypot(*mid) > 0:
                mid = mid * (ra / np.hypot(*mid))
                if pa[0] == 0.0 and pb[0] == 0.0:
                    mid[0] = 0.0
            crd[8 * e + 2 * k + 1] = mid
    lnods = np.arange(1, 8 * nelem + 1, dtype=np.int32).reshape(nelem, 8)
    ax_el = np.array([e + 1 for e, c in enumerate(corners) if c[0][0] == 0.0 and c[3][0] == 0.0], dtype=np.int32)
    for e, c in enumerate(corners):                 # no element touches the axis with a corner only
        n_ax = sum(s_ == 0.0 for s_, _ in c)
        assert n_ax in (0, 2) and (n_ax == 0 or e + 1 in ax_el), (e, c)
    # ---- solid / fluid domains: global element order = solid elements first; numbers per domain; the
    # boundary pairs with the j index of the shared edge on either side (data_mesh.f90:106-110)
    dom = {}
    if fluid is not None:
        rmid = np.array([np.hypot(np.mean([p[0] for p in c]), np.mean([p[1] for p in c])) for c in corners])
        is_f = np.array([bool(fluid(r)) for r in rmid])
        perm = np.concatenate([np.nonzero(~is_f)[0], np.nonzero(is_f)[0]])
        ns, nf = int((~is_f).sum()), int(is_f.sum())
        corners = [corners[e] for e in perm]
        eltype = [eltype[e] for e in perm]
        coarsing = [coarsing[e] for e in perm]
        crd = crd.reshape(nelem, 8, 2)[perm].reshape(-1, 2)
        igr = ig[perm]

        def compress(block):
            flat_ = block.reshape(-1)
            _, first_, inv_ = np.unique(flat_, return_index=True, return_inverse=True)
            order_ = np.argsort(np.argsort(first_))
            return (order_[inv_] + 1).astype(np.int32)

        ig_s, ig_f = compress(igr[:ns]), compress(igr[ns:])
        edge_users = {}
        for e in range(nelem):
            for jj, pts in ((0, igr[e, 0, :]), (4, igr[e, 4, :]), (-1, igr[e, :, 0]), (-2, igr[e, :, 4])):
                edge_users.setdefault(tuple(sorted((int(pts[0]), int(pts[4])))), []).append((e, jj, [int(q) for q in pts]))
        bs, bf, js, jf = [], [], [], []
        for users in edge_users.values():
            if len(users) == 2 and (users[0][0] < ns) != (users[1][0] < ns):
                (es, j_s, ps), (ef, j_f, pf) = sorted(users)
                assert j_s in (0, 4) and j_f in (0, 4) and ps == pf, "a solid/fluid boundary must be an eta = const edge with xi aligned"
                bs.append(es + 1); bf.append(ef - ns + 1); js.append(j_s); jf.append(j_f)
        ax_all = np.array([e + 1 for e, c in enumerate(corners) if c[0][0] == 0.0 and c[3][0] == 0.0], dtype=np.int32)
        ig = igr
        dom = dict(nel_solid=ns, nel_fluid=nf, igloc_solid=ig_s, igloc_fluid=ig_f, nglob_solid=int(ig_s.max()),
                   nglob_fluid=int(ig_f.max()) if nf else 0, bdry_solid_el=np.array(bs, np.int32), bdry_fluid_el=np.array(bf, np.int32),
                   bdry_jpol_solid=np.array(js, np.int32), bdry_jpol_fluid=np.array(jf, np.int32),
                   ax_el_solid=ax_all[ax_all <= ns], ax_el_fluid=ax_all[ax_all > ns] - ns)
        ax_el = ax_all
    return dict(nelem=nelem, crd=crd, lnods=lnods, eltype=eltype, coarsing=np.array(coarsing), igloc=igloc, nglob=nglob,
                ax_el=ax_el, corners=corners, router=router, rmin=rmin, ndoubling=int(np.sum(coarsing)), topo=ig, **dom)


def partition(M, nth_blocks, r_cuts=()):
    """The mesher's domain decomposition of a mesh built above: theta blocks by the colatitude of the element
    centroids, times radial blocks cut at `r_cuts` (never at a solid/fluid boundary) — one database per rank,
    with its own element order (solid first), its own global numbers per domain, and the message lists of
    data_comm.f90:36-71 towards every rank it shares points with (the same points in the same order on either
    side; a point shared by three or four ranks travels in each pairwise message)."""
    nelem = M["nelem"]
    ns = M.get("nel_solid", nelem)
    cent = np.array([[np.mean([p[0] for p in c]), np.mean([p[1] for p in c])] for c in M["corners"]])
    th = np.arctan2(cent[:, 0], cent[:, 1])
    rr = np.hypot(cent[:, 0], cent[:, 1])
    tb = np.minimum((th / np.pi * nth_blocks).astype(int), nth_blocks - 1)
    rb = np.searchsorted(np.asarray(r_cuts, dtype=float), rr)
    nrb = len(r_cuts) + 1
    rank = tb * nrb + rb
    if "bdry_solid_el" in M:                      # a boundary pair lives on one rank
        for es, ef in zip(M["bdry_solid_el"], M["bdry_fluid_el"]):
            rank[ns + ef - 1] = rank[es - 1]
    nranks = nth_blocks * nrb
    topo = M["topo"].reshape(nelem, 25)
    is_f = np.arange(nelem) >= ns
    out, gids = [], []
    for r in range(nranks):
        E = np.nonzero(rank == r)[0]              # global order is solid first: so is this
        Es, Ef = E[~is_f[E]], E[is_f[E]]
        loc = {}

        def compress(block):
            flat_ = block.reshape(-1)
            if flat_.size == 0:
                return np.zeros(0, np.int32), np.zeros(0, np.int64)
            u, first_, inv_ = np.unique(flat_, return_index=True, return_inverse=True)
            order_ = np.argsort(np.argsort(first_))
            gid = np.zeros(u.size, np.int64)
            gid[order_] = u                       # gid[local - 1] = topological id
            return (order_[inv_] + 1).astype(np.int32), gid

        ig_s, gid_s = compress(topo[Es])
        ig_f, gid_f = compress(topo[Ef])
        gids.append((gid_s, gid_f))
        new_of = {int(e): k for k, e in enumerate(E)}
        ne = E.size
        d = dict(nelem=ne, nel_solid=int(Es.size), nel_fluid=int(Ef.size), igloc_solid=ig_s, igloc_fluid=ig_f,
                 nglob_solid=int(gid_s.size), nglob_fluid=int(gid_f.size), igloc=ig_s, nglob=int(gid_s.size),
                 crd=M["crd"].reshape(nelem, 8, 2)[E].reshape(-1, 2), lnods=np.arange(1, 8 * ne + 1, dtype=np.int32).reshape(ne, 8),
                 eltype=[M["eltype"][e] for e in E], coarsing=np.asarray(M["coarsing"])[E], router=M["router"], rmin=M["rmin"],
                 nranks=nranks, rank=r, have_fluid=M.get("nel_fluid", 0) > 0)
        ax = np.array([new_of[int(e) - 1] + 1 for e in M["ax_el"] if int(e) - 1 in new_of], dtype=np.int32)
        d["ax_el"] = ax
        d["ax_el_solid"] = ax[ax <= Es.size]
        d["ax_el_fluid"] = ax[ax > Es.size] - Es.size
        bs, bf, js, jf = [], [], [], []
        if "bdry_solid_el" in M:
            for es, ef, a_, b_ in zip(M["bdry_solid_el"], M["bdry_fluid_el"], M["bdry_jpol_solid"], M["bdry_jpol_fluid"]):
                if int(es) - 1 in new_of:
                    assert ns + int(ef) - 1 in new_of
                    bs.append(new_of[int(es) - 1] + 1); bf.append(new_of[ns + int(ef) - 1] - Es.size + 1); js.append(a_); jf.append(b_)
        d.update(bdry_solid_el=np.array(bs, np.int32), bdry_fluid_el=np.array(bf, np.int32),
                 bdry_jpol_solid=np.array(js, np.int32), bdry_jpol_fluid=np.array(jf, np.int32))
        out.append(d)
    for r in range(nranks):
        for dname, k in (("solid", 0), ("fluid", 1)):
            peers, lists = [], []
            mine = gids[r][k]
            where = {int(g): i + 1 for i, g in enumerate(mine)}
            for q in range(nranks):
                if q == r:
                    continue
                shared = np.intersect1d(mine, gids[q][k])
                if shared.size:
                    peers.append(q)
                    lists.append(np.array([where[int(g)] for g in shared], dtype=np.int32))
            out[r]["halo_" + dname] = (np.array(peers, np.int32), lists)
    return out


def write_database(path, M, basis, *, bkgrdmodel="prem_iso_light", dt=1.0, period=50.0,
                   discont=(6371e3, 6291e3, 6151e3, 5971e3, 5771e3, 5701e3, 5600e3, 3630e3, 3480e3, 1221.5e3),
                   solid_domain=None):
    """One rank: the record sequence of MESHER/pdb.f90:2205-2382 (see axisem_b200/host/meshdb_io.py)."""
    f = open(path, "wb")

    def rec(*parts):
        payload = b"".join(parts)
        m = struct.pack("<i", len(payload))
        f.write(m + payload + m)

    I = lambda *v: np.asarray(v, dtype="<i4").tobytes()
    D = lambda *v: np.asarray(v, dtype="<f8").tobytes()
    A = lambda a, dt_: np.ascontiguousarray(a, dtype=dt_).tobytes()
    nelem, npol = M["nelem"], 4
    ns, nf = M.get("nel_solid", nelem), M.get("nel_fluid", 0)
    ig_s, ig_f = M.get("igloc_solid", M["igloc"]), M.get("igloc_fluid", np.zeros(0, np.int32))
    ng_s, ng_f = M.get("nglob_solid", M["nglob"]), M.get("nglob_fluid", 0)
    nb = len(M.get("bdry_solid_el", ()))
    for v in (M.get("nranks", 1), npol, nelem, nelem * 25, ns, nf, ns * 25, nf * 25, ng_s, ng_f, nb, len(discont), len(bkgrdmodel)):
        rec(I(v))
    for name in ("xi_k", "eta", "dxi", "wt", "wt_axial_k"):
        rec(A(getattr(basis, name, np.zeros(npol + 1)), "<f8"))
    rec(A(basis.G0, "<f4"))
    for name in ("G1", "G1T", "G2", "G2T"):
        rec(A(np.asarray(getattr(basis, name), dtype=np.float32).T, "<f4"))
    rec(I(M["crd"].shape[0]))
    rec(A(M["crd"][:, 0], "<f8"))
    rec(A(M["crd"][:, 1], "<f8"))
    for e in range(nelem):
        rec(A(M["lnods"][e], "<i4"))
    rec(I(ng_s + ng_f))
    rec(b"".join(t.encode().ljust(6)[:6] for t in M["eltype"]))
    rec(A(np.asarray(M["coarsing"]).astype(np.int32), "<i4"))
    rec(A(np.arange(1, ns + 1), "<i4"))                       # ielsolid
    rec(A(np.arange(ns + 1, nelem + 1), "<i4"))               # ielfluid
    rec(A(ig_s, "<i4"))
    rec(A(ig_f, "<i4"))
    rec(I(1 if nb else 0))                                    # have_bdry_elem
    if nb:
        for k in ("bdry_solid_el", "bdry_fluid_el", "bdry_jpol_solid", "bdry_jpol_fluid"):
            rec(A(M[k], "<i4"))
    rec(D(1.5, period, 0.6, dt))
    rec(bkgrdmodel.encode())
    rec(b"none  ")
    have_fluid = bool(M.get("have_fluid", nf > 0))            # of the whole mesh
    rec(D(M["router"]), I(1 if have_fluid else 0))            # router, have_fluid
    if solid_domain is None:
        solid_domain = [1] * len(discont)
    for r, sd in zip(discont, solid_domain):
        rec(D(r), I(sd), I(0))
    rec(D(M["rmin"], 0.0, 0.0, 0.0))
    rec(D(0.0, 0.0))
    rec(D(0.0, 0.0))
    for _ in range(2):
        rec(D(0.0), I(1))
        rec(D(0.0, 0.0))
    ax = M["ax_el"]
    ax_s, ax_f = M.get("ax_el_solid", ax), M.get("ax_el_fluid", np.zeros(0, np.int32))
    rec(I(ax.size, ax_s.size, ax_f.size))
    rec(A(ax, "<i4"))
    rec(A(ax_s, "<i4"))
    rec(A(ax_f, "<i4"))
    for dname in ("solid", "fluid"):                          # messaging (pdb.f90:2330-2380)
        if dname == "fluid" and not have_fluid:
            break
        peers, lists = M.get("halo_" + dname, (np.zeros(0, np.int32), []))
        rec(I(len(peers)))
        if len(peers):
            rec(A(peers, "<i4"))
            rec(A([len(l) for l in lists], "<i4"))
            for l in lists:
                for g in l:
                    rec(I(int(g)))
    f.close()
