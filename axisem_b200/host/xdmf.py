"""XDMF snapshots: the plot-point maps of `dump_xdmf_grid` (SOLVER/meshes_io.F90:110-437) and the
files `glob_snapshot_xdmf` / `finish_xdmf_xml` write (SOLVER/wavefields_io.f90:195-598,
meshes_io.F90:441-464) — `xdmf_points_NNNN.dat`, `xdmf_grid_NNNN.dat`,
`xdmf_snap_{s,p,z,trace,curlip}_NNNN.dat` (big-endian stream files), `xdmf_meshonly_NNNN.xdmf`
and `xdmf_xml_NNNN.xdmf`.

The fields themselves are formed by the time loop (axb_set_xdmf / axb_fetch_xdmf); this module is
the host on either side of it."""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import numpy as np

from .mesh import LocalMesh


def xdmf_maps(mesh: LocalMesh, i_arr: Sequence[int] = (0, 2, 4), j_arr: Sequence[int] = (0, 2, 4),
              rmin: float = 0.0, rmax: float = 7.0e6, thetamin: float = 0.0, thetamax: float = np.pi) -> Dict:
    """dump_xdmf_grid: elements inside [rmin, rmax] x [thetamin, thetamax] (any corner test of
    :126-160), plot points = the (i_arr, j_arr) GLL points of those elements de-duplicated by
    global number, first visit wins in the order fluid elements, solid elements; i outer, j inner.
    Defaults as parameters.F90:428-431 (XDMF_GLL_I / _J have no default there; 0 2 4 is what the
    reference's inparam_advanced template sets).

    Returns plotting_mask, mapping_ijel_iplot as (nelem, j_n, i_n) int32 arrays (= Fortran
    (i_n, j_n, nelem)), fluid elements first; npoint_plot; nelem_plot; points (npoint_plot, 2) =
    (s, z); grid (nelem_plot, 4) 0-based corner numbers."""
    i_arr = np.asarray(i_arr, dtype=np.int32)
    j_arr = np.asarray(j_arr, dtype=np.int32)
    i_n, j_n = i_arr.size, j_arr.size
    nf, ns = mesh.nel_fluid, mesh.nel_solid
    nelem = nf + ns
    mask = np.zeros((nelem, j_n, i_n), dtype=np.int32)
    mapping = np.zeros((nelem, j_n, i_n), dtype=np.int32)
    in_range = np.zeros(nelem, dtype=bool)
    coords = {}
    for dom, off, nel in (("fluid", 0, nf), ("solid", nf, ns)):
        if nel == 0:
            continue
        _, th, r, s, z, _, _ = mesh.coords(dom)
        coords[dom] = (s, z)
        # corners (0,0), (0,npol), (npol,0), (npol,npol); r depends on j only, theta on i only
        rc = np.stack([r[:, 0], r[:, 4]], axis=1)
        tc = np.stack([th[:, 0], th[:, 4]], axis=1)
        in_range[off:off + nel] = ((rc.min(axis=1) < rmax) & (rc.max(axis=1) > rmin)
                                   & (tc.min(axis=1) < thetamax) & (tc.max(axis=1) > thetamin))
    nelem_plot = int(in_range.sum()) * (i_n - 1) * (j_n - 1)
    ct = 0
    seen: Dict[int, int] = {}
    for dom, off, nel, ig, gofs in (("fluid", 0, nf, mesh.igloc_fluid, 0),
                                    ("solid", nf, ns, mesh.igloc_solid, mesh.nglob_fluid)):
        if nel == 0:
            continue
        ig = np.asarray(ig).reshape(nel, 5, 5)                  # [el][jpol][ipol]
        for el in range(nel):
            if not in_range[off + el]:
                continue
            for i in range(i_n):
                for j in range(j_n):
                    idest = int(ig[el, j_arr[j], i_arr[i]]) + gofs
                    if idest not in seen:
                        ct += 1
                        seen[idest] = ct
                        mask[off + el, j, i] = 1
                    mapping[off + el, j, i] = seen[idest]
    npoint_plot = ct
    points = np.zeros((npoint_plot, 2), dtype=np.float32)
    for dom, off, nel in (("fluid", 0, nf), ("solid", nf, ns)):
        if nel == 0:
            continue
        s, z = coords[dom]
        el, j, i = np.nonzero(mask[off:off + nel])
        k = mapping[off:off + nel][el, j, i] - 1
        points[k, 0] = s[el, j_arr[j], i_arr[i]]
        points[k, 1] = z[el, j_arr[j], i_arr[i]]
    grid = np.zeros((nelem_plot, 4), dtype=np.int32)
    c = 0
    for el in np.nonzero(in_range)[0]:
        for i in range(i_n - 1):
            for j in range(j_n - 1):
                grid[c] = (mapping[el, j, i] - 1, mapping[el, j, i + 1] - 1,
                           mapping[el, j + 1, i + 1] - 1, mapping[el, j + 1, i] - 1)
                c += 1
    return {"i_arr": i_arr, "j_arr": j_arr, "plotting_mask": mask, "mapping_ijel_iplot": mapping,
            "npoint_plot": npoint_plot, "nelem_plot": nelem_plot, "points": points, "grid": grid}


def _hyperslab(name: str, npoint: int, isnap0: int, nsnap: int, fname: str) -> str:
    # the Attribute block of formats 734 / 735 (wavefields_io.f90:286-297)
    return (f'        <Attribute Name="{name}" AttributeType="Scalar" Center="Node">\n'
            f'            <DataItem ItemType="HyperSlab" Dimensions="{npoint:10d}" Type="HyperSlab">\n'
            f'                <DataItem Dimensions="3 2" Format="XML">\n'
            f'                    {isnap0:10d}          0 \n'
            f'                             1          1 \n'
            f'                             1 {npoint:10d}\n'
            f'                </DataItem>\n'
            f'                <DataItem Dimensions="{nsnap:10d}{npoint:10d}" NumberType="Float" Format="binary" Endian="Big">\n'
            f'                   {fname}\n'
            f'                </DataItem>\n'
            f'            </DataItem>\n'
            f'        </Attribute>\n')


def _abs_block(npoint: int, app: str, comps: Sequence[str]) -> str:
    terms = " + ".join(f"${k} * ${k}" for k in range(len(comps)))
    refs = "".join(
        f'                <DataItem Reference="XML">\n'
        f'                    /Xdmf/Domain/Grid[@Name="CellsTime"]/Grid[@Name="{app}"]/Attribute[@Name="{c}"]/DataItem[1]\n'
        f'                </DataItem>\n' for c in comps)
    return (f'        <Attribute Name="abs" AttributeType="Scalar" Center="Node">\n'
            f'            <DataItem ItemType="Function" Function="sqrt({terms})" Dimensions="{npoint:10d}">\n'
            f'{refs}'
            f'            </DataItem>\n'
            f'        </Attribute>\n')


def snapshot_xml(isnap: int, t: float, maps: Dict, nsnap: int, appmynum: str, monopole: bool) -> str:
    """One `<Grid>` of xdmf_xml_NNNN.xdmf: formats 734 (monopole) / 735 of glob_snapshot_xdmf."""
    app = f"{isnap:04d}"
    npnt, nel = maps["npoint_plot"], maps["nelem_plot"]
    comps = ("u_s", "u_z") if monopole else ("u_s", "u_p", "u_z")
    files = {"u_s": "s", "u_p": "p", "u_z": "z"}
    out = (f'    <Grid Name="{app}" GridType="Uniform">\n'
           f'        <Time Value="{t:8.2f}" />\n'
           f'        <Topology TopologyType="Quadrilateral" NumberOfElements="{nel:10d}">\n'
           f'            <DataItem Reference="XML">\n'
           f'                /Xdmf/Domain/DataItem[@Name="grid"]\n'
           f'            </DataItem>\n'
           f'        </Topology>\n'
           f'        <Geometry GeometryType="XY">\n'
           f'            <DataItem Reference="XML">\n'
           f'                /Xdmf/Domain/DataItem[@Name="points"]\n'
           f'            </DataItem>\n'
           f'        </Geometry>\n')
    for c in comps:
        out += _hyperslab(c, npnt, isnap - 1, nsnap, f"xdmf_snap_{files[c]}_{appmynum}.dat")
    out += _abs_block(npnt, app, comps)
    out += _hyperslab("straintrace", npnt, isnap - 1, nsnap, f"xdmf_snap_trace_{appmynum}.dat")
    out += _hyperslab("curlinplane", npnt, isnap - 1, nsnap, f"xdmf_snap_curlip_{appmynum}.dat")
    out += '    </Grid>\n\n'
    return out


def write_xdmf(outdir: str, rank: int, maps: Dict, fields: np.ndarray, times: Sequence[float], *,
               monopole: bool, nsnap_total: Optional[int] = None) -> Dict[str, str]:
    """Write everything the reference leaves in Data/ for one rank.  `fields`: (5, nsnap,
    npoint_plot) as TimeLoop.xdmf_snapshots(): u_s, u_p, u_z, straintrace, curlinplane."""
    os.makedirs(outdir, exist_ok=True)
    app = f"{rank:04d}"
    nsnap = int(fields.shape[1])
    nsnap_total = nsnap if nsnap_total is None else int(nsnap_total)
    npnt, nel = maps["npoint_plot"], maps["nelem_plot"]
    paths = {}

    def put(name, arr, dt):
        paths[name] = os.path.join(outdir, name)
        np.ascontiguousarray(arr).astype(dt).tofile(paths[name])

    put(f"xdmf_points_{app}.dat", maps["points"], ">f4")       # points(1:2, npoint_plot)
    put(f"xdmf_grid_{app}.dat", maps["grid"], ">i4")           # grid(1:4, nelem_plot)
    names = ["s", "p", "z", "trace", "curlip"]
    for v, n in enumerate(names):
        if n == "p" and monopole:
            continue                                              # unit 13101 is not opened
        put(f"xdmf_snap_{n}_{app}.dat", fields[v], ">f4")       # one record per snapshot
    head = ('<?xml version="1.0" ?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n'
            '<Xdmf xmlns:xi="http://www.w3.org/2003/XInclude" Version="2.2">\n<Domain>\n')
    mesh_only = (head +
                 '<Grid Name="CellsTime" GridType="Collection" CollectionType="Temporal">\n'
                 '  <Grid GridType="Uniform">\n    <Time Value="0.000" />\n'
                 f'    <Topology TopologyType="Quadrilateral" NumberOfElements="{nel:10d}">\n'
                 f'      <DataItem Dimensions="{nel:10d} 4" NumberType="Int" Format="binary" Endian="Big">\n'
                 f'        xdmf_grid_{app}.dat\n      </DataItem>\n    </Topology>\n'
                 '    <Geometry GeometryType="XY">\n'
                 f'      <DataItem Dimensions="{npnt:10d} 2" NumberType="Float" Format="binary" Endian="Big">\n'
                 f'        xdmf_points_{app}.dat\n      </DataItem>\n    </Geometry>\n'
                 '  </Grid>\n</Grid>\n</Domain>\n</Xdmf>\n')
    paths["meshonly"] = os.path.join(outdir, f"xdmf_meshonly_{app}.xdmf")
    open(paths["meshonly"], "w").write(mesh_only)
    xml = (head + '\n'
           f'<DataItem Name="grid" Dimensions="{nel:10d} 4" NumberType="Int" Format="binary" Endian="Big">\n'
           f'  xdmf_grid_{app}.dat\n</DataItem>\n'
           f'<DataItem Name="points" Dimensions="{npnt:10d} 2" NumberType="Float" Format="binary" Endian="Big">\n'
           f'  xdmf_points_{app}.dat\n</DataItem>\n\n'
           '<Grid Name="CellsTime" GridType="Collection" CollectionType="Temporal">\n\n')
    for k in range(nsnap):
        xml += snapshot_xml(k + 1, float(times[k]), maps, nsnap_total, app, monopole)
    xml += '</Grid>\n</Domain>\n</Xdmf>\n'                      # finish_xdmf_xml
    paths["xml"] = os.path.join(outdir, f"xdmf_xml_{app}.xdmf")
    open(paths["xml"], "w").write(xml)
    return paths
