"""Write one rank's mesh in the MESHER's `meshdb.datNNNN` format (test-fixture generator).

Record for record the stream of MESHER/pdb.f90:2205-2382 (`write_db`): Fortran sequential
unformatted, 4-byte record markers, default integer/logical = 4 bytes, `dp` = real(8), the
derivative matrices in the mesher's `realkind`.  The reader is native code
(axisem_b200/hostcxx/meshdb.cpp, the counterpart of SOLVER/data_mesh.f90:190-322 +
get_mesh.f90:101-383); this writer only exists so that the reader can be exercised without
the Fortran mesher, on the synthetic meshes of this package.
"""
from __future__ import annotations

import struct

import numpy as np


class _Unformatted:
    def __init__(self, path):
        self.f = open(path, "wb")

    def rec(self, *parts):
        payload = b"".join(parts)
        m = struct.pack("<i", len(payload))
        self.f.write(m + payload + m)

    def close(self):
        self.f.close()


def _i(*v):
    return np.asarray(v, dtype="<i4").tobytes()


def _d(*v):
    return np.asarray(v, dtype="<f8").tobytes()


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt).tobytes()


def control_nodes(mesh):
    """crd_nodes (npoin,2) [m] and lnods (nelem,8): 8 control nodes per element (corners and
    edge mid-points, counter-clockwise from (xi,eta)=(-1,-1)); not shared between elements."""
    th_a = np.concatenate([mesh.solid.th_a, mesh.fluid.th_a])
    th_b = np.concatenate([mesh.solid.th_b, mesh.fluid.th_b])
    r_a = np.concatenate([mesh.solid.r_a, mesh.fluid.r_a])
    r_b = np.concatenate([mesh.solid.r_b, mesh.fluid.r_b])
    th_m, r_m = 0.5 * (th_a + th_b), 0.5 * (r_a + r_b)
    th = np.stack([th_a, th_m, th_b, th_b, th_b, th_m, th_a, th_a], axis=1)
    r = np.stack([r_a, r_a, r_a, r_m, r_b, r_b, r_b, r_m], axis=1)
    s, z = r * np.sin(th), r * np.cos(th)
    nelem = th_a.size
    lnods = np.arange(1, 8 * nelem + 1, dtype=np.int32).reshape(nelem, 8)
    return np.stack([s.reshape(-1), z.reshape(-1)], axis=1), lnods


def write_meshdb(mesh, path: str, *, period: float = 50.0, courant: float = 0.6, dt: float = 0.1,
                 bkgrdmodel: str = "prem_iso", eltype=None):
    b, spec = mesh.basis, mesh.spec
    npol = spec.npol
    ns, nf = mesh.nel_solid, mesh.nel_fluid
    nelem = ns + nf
    n2 = (npol + 1) ** 2
    router = float(spec.router)
    discs = [float(l.r_top) for l in spec.layers][::-1]
    ndisc = len(discs)
    u = _Unformatted(path)
    for v in (mesh.nranks, npol, nelem, nelem * n2, ns, nf, ns * n2, nf * n2, mesh.nglob_solid,
              mesh.nglob_fluid, mesh.nel_bdry, ndisc, len(bkgrdmodel)):
        u.rec(_i(v))
    # spectral stuff: xi_k, eta, dxi, wt, wt_axial_k (dp); G0, G1, G1T, G2, G2T (realkind = sp)
    for name in ("xi_k", "eta", "dxi", "wt", "wt_axial_k"):
        u.rec(_arr(getattr(b, name, np.zeros(npol + 1)), "<f8"))     # dxi is not used by the SOLVER loop
    u.rec(_arr(b.G0, "<f4"))
    for name in ("G1", "G1T", "G2", "G2T"):
        u.rec(_arr(np.asarray(getattr(b, name), dtype=np.float32).T, "<f4"))     # column-major
    crd, lnods = control_nodes(mesh)
    u.rec(_i(crd.shape[0]))
    u.rec(_arr(crd[:, 0], "<f8"))
    u.rec(_arr(crd[:, 1], "<f8"))
    for e in range(nelem):
        u.rec(_arr(lnods[e], "<i4"))
    u.rec(_i(mesh.nglob_solid + mesh.nglob_fluid))              # nglob of the rank
    if eltype is None:
        u.rec(b"".join(b"curved" for _ in range(nelem)))        # eltype, character(len=6)
    else:
        u.rec(b"".join(bytes(t).ljust(6)[:6] for t in eltype))  # (solid elements first, then fluid)
    u.rec(_arr(np.zeros(nelem), "<i4"))                         # coarsing, logical
    u.rec(_arr(np.arange(1, ns + 1), "<i4"))                    # ielsolid
    u.rec(_arr(np.arange(ns + 1, nelem + 1), "<i4"))            # ielfluid
    u.rec(_arr(mesh.igloc_solid, "<i4"))
    u.rec(_arr(mesh.igloc_fluid, "<i4"))
    u.rec(_i(1 if mesh.nel_bdry else 0))
    if mesh.nel_bdry:
        for k in ("bdry_solid_el", "bdry_fluid_el", "bdry_jpol_solid", "bdry_jpol_fluid"):
            u.rec(_arr(getattr(mesh, k), "<i4"))
    u.rec(_d(1.5, period, courant, dt))                          # pts_wavelngth, period, courant, dt
    u.rec(bkgrdmodel.encode())
    u.rec(b"none  ")                                             # override_ext_q, character(len=6)
    u.rec(_d(router), _i(1 if nf else 0))
    solid_dom = [0 if l.fluid else 1 for l in spec.layers][::-1]
    for k in range(ndisc):
        u.rec(_d(discs[k]), _i(solid_dom[k]), _i(0))
    u.rec(_d(float(spec.layers[0].r_bot), 0.0, 0.0, 0.0))        # rmin, minh_ic, maxh_ic, maxh_icb
    u.rec(_d(0.0, 0.0))
    u.rec(_d(0.0, 0.0))
    for _ in range(2):
        u.rec(_d(0.0), _i(1))
        u.rec(_d(0.0, 0.0))
    ax_s, ax_f = mesh.ax_el_solid, mesh.ax_el_fluid
    ax = np.concatenate([ax_s, ns + ax_f]) if nelem else np.zeros(0, np.int32)
    u.rec(_i(ax.size, ax_s.size, ax_f.size))
    u.rec(_arr(ax, "<i4"))
    u.rec(_arr(ax_s, "<i4"))
    u.rec(_arr(ax_f, "<i4"))
    for dom, hs in (("solid", mesh.halo_solid), ("fluid", mesh.halo_fluid)):
        if dom == "fluid" and nf == 0:
            break
        u.rec(_i(hs.nmsg))
        if hs.nmsg:
            u.rec(_arr(hs.list_peer, "<i4"))
            u.rec(_arr(hs.sizemsg, "<i4"))
            for m in range(hs.nmsg):
                for ip in range(int(hs.sizemsg[m])):
                    u.rec(_i(int(hs.glocal_index_msg[m, ip])))
    u.close()


def read_axbprob(path: str):
    """{name: ndarray} of an AXBPROB1 container (problem_bin.py)."""
    out = {}
    dts = {0: "<f4", 1: "<f8", 2: "<i4"}
    with open(path, "rb") as f:
        assert f.read(8) == b"AXBPROB1"
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (nl,) = struct.unpack("<H", f.read(2))
            name = f.read(nl).decode()
            t, nd = struct.unpack("<BB", f.read(2))
            dims = struct.unpack("<%dQ" % nd, f.read(8 * nd)) if nd else ()
            (nb,) = struct.unpack("<Q", f.read(8))
            out[name] = np.frombuffer(f.read(nb), dtype=dts[t]).reshape(dims)
    return out
