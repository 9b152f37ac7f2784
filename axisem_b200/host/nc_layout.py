"""The output database of the solver in the layout of `axisem_output.nc4` — what Instaseis and the
kernel code read (SOLVER/nc_routines.F90:829-1492 `nc_define_outputfile`, global attributes of
parameters.F90:1480-1552, mesh variables of meshes_io.F90:489-850).

This image has neither libnetcdf nor libhdf5 (and no h5py / netCDF4), so the database is written as
a *directory*: `schema.json` — groups, dimensions, variables (type, dimension names in netCDF /
C order, i.e. the reverse of the Fortran `dimids`, chunking, attributes) and global attributes,
exactly the definitions of `nc_define_outputfile` — plus one raw little-endian file per
variable, `<Group>/<variable>.bin`, in that dimension order.  `tools/pack_netcdf.py` turns the
directory into the NetCDF-4 file wherever netCDF4 (or h5py) exists; nothing is lost or renamed
on the way.  tests/test_nc_layout.py holds the schema against a fixture extracted from the
reference's Fortran (tests/golden/make_nc_schema_fixture.py).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from .source import stf_shift

SNAP_VARS = {   # nc_varnamelist, nc_routines.F90:943-1045
    ("displ_only", True): ["disp_s", "disp_z"],
    ("displ_only", False): ["disp_s", "disp_p", "disp_z"],
    ("strain_only", True): ["strain_dsus", "strain_dsuz", "strain_dpup", "straintrace"],
    ("strain_only", False): ["strain_dsus", "strain_dsuz", "strain_dpup", "strain_dsup", "strain_dzup", "straintrace"],
    ("fullfields", True): ["strain_dsus", "strain_dsuz", "strain_dpup", "straintrace", "velo_s", "velo_z"],
    ("fullfields", False): ["strain_dsus", "strain_dsuz", "strain_dpup", "strain_dsup", "strain_dzup", "straintrace",
                            "velo_s", "velo_p", "velo_z"],
}


def _var(dtype: str, dims: Sequence[str], **kw) -> Dict:
    d = {"dtype": dtype, "dims": list(dims)}
    d.update(kw)
    return d


def schema(*, nrec: int, nseismo: int, niter: int, dump_wavefields: bool, dump_type: str = "displ_only",
           monopole: bool = False, npoints_global: int = 0, nstrain: int = 0, nelem_kwf_global: int = 0,
           anel: bool = False, npol: int = 4, ibeg: int = 0, iend: int = 4, jbeg: int = 0, jend: int = 4) -> Dict:
    """Groups / dimensions / variables as `nc_define_outputfile` defines them.  Dimension lists
    are in netCDF (C) order; the Fortran `dimids=[a, b, c]` of the reference reads [c, b, a] here."""
    root_dims = {"seis_timesteps": nseismo, "sim_timesteps": niter, "components": 3}
    seis_dims, seis_vars = {}, {}
    if nrec > 0:
        seis_dims = {"recnamlength": 40, "receivers": nrec}
        seis_vars.update({
            "displacement": _var("f4", ["receivers", "components", "seis_timesteps"], chunks=[1, 3, nseismo],
                                 attrs={"units": "meters", "_FillValue": 0.0}),
            "time": _var("f8", ["seis_timesteps"]),
            "phi": _var("f4", ["receivers"]),
            "theta_requested": _var("f4", ["receivers"]),
            "theta": _var("f4", ["receivers"]),
            "processor_of_receiver": _var("i4", ["receivers"]),
            "receiver_name": _var("S1", ["recnamlength", "receivers"]),
        })
    for n in ("stf_seis", "stf_d_seis"):
        seis_vars[n] = _var("f4", ["seis_timesteps"])
    for n in ("stf_iter", "stf_d_iter"):
        seis_vars[n] = _var("f4", ["sim_timesteps"])
    groups = {"Seismograms": {"dimensions": seis_dims, "variables": seis_vars},
              "Snapshots": {"dimensions": {}, "variables": {}},
              "Surface": {"dimensions": {}, "variables": {}},
              "Mesh": {"dimensions": {}, "variables": {}}}
    root_vars = {}
    if dump_wavefields:
        root_dims.update({"snapshots": nstrain, "gllpoints_all": npoints_global})
        root_vars["snapshot_times"] = _var("f4", ["snapshots"])
        mesh = groups["Mesh"]
        if dump_type == "displ_only":
            mesh["dimensions"].update({"elements": nelem_kwf_global, "control_points": 4, "npol": npol + 1})
        for n in ("mesh_S", "mesh_Z"):
            mesh["variables"][n] = _var("f8", ["gllpoints_all"])
        for n in ("mesh_vp", "mesh_vs", "mesh_rho", "mesh_lambda", "mesh_mu", "mesh_xi", "mesh_phi", "mesh_eta"):
            mesh["variables"][n] = _var("f4", ["gllpoints_all"])
        if anel:
            for n in ("mesh_Qmu", "mesh_Qka"):
                mesh["variables"][n] = _var("f4", ["gllpoints_all"])
        if dump_type == "displ_only":
            for n in ("midpoint_mesh", "eltype", "axis"):
                mesh["variables"][n] = _var("i4", ["elements"])
            mesh["variables"]["fem_mesh"] = _var("i4", ["elements", "control_points"])
            mesh["variables"]["sem_mesh"] = _var("i4", ["elements", "npol", "npol"])
            for n in ("mp_mesh_S", "mp_mesh_Z"):
                mesh["variables"][n] = _var("f8", ["elements"])
            mesh["variables"]["G0"] = _var("f8", ["npol"])
            for n in ("G1", "G2"):
                mesh["variables"][n] = _var("f8", ["npol", "npol"])
            for n in ("gll", "glj"):
                mesh["variables"][n] = _var("f8", ["npol"])
        snap = groups["Snapshots"]["variables"]
        for n in SNAP_VARS[(dump_type, monopole)]:
            snap[n] = _var("f4", ["snapshots", "gllpoints_all"])
        for n in ("stf_dump", "stf_d_dump"):
            snap[n] = _var("f4", ["snapshots"])
    return {"format": "axisem_output.nc4 (NetCDF-4), written as a directory", "dimensions": root_dims,
            "variables": root_vars, "groups": groups, "attributes": {}}


def global_attributes(prob, *, nseismo: int, nstrain: int, deltat_coarse: float, num_rec_tot: int,
                      dump_wavefields: bool, dump_type: str, background_model: str = "prem_iso",
                      srccolat: float = 0.0, srclon: float = 0.0, simtype: str = "single") -> Dict:
    """parameters.F90:1480-1552 (build provenance strings are this repository's)."""
    s = prob.source
    shift = stf_shift(s, prob.deltat)
    seis_dt = prob.deltat * prob.seis_it
    spec = prob.mesh.spec
    a = {
        "file version": 10, "background model": background_model, "external model name": "",
        "attenuation": int(bool(prob.anel)), "planet radius": spec.router / 1000.0,
        "datetime": "", "git commit hash": "", "user name": "", "host name": "",
        "compiler brand": "nvcc", "compiler version": "", "FFLAGS": "", "CFLAGS": "", "LDFLAGS": "", "OpenMP": "no",
        "time scheme": prob.time_scheme, "time step in sec": np.float32(prob.deltat).item(),
        "number of time steps": int(prob.niter), "npol": 4,
        "excitation type": prob.src_type, "source type": s.src_type2, "source time function": s.stf_type,
        "simulation type": simtype, "dominant source period": np.float32(s.t_0).item(),
        "source depth in km": np.float32(s.depth / 1000.0).item(),
        "Source colatitude": np.float32(srccolat).item(), "Source longitude": np.float32(srclon).item(),
        "scalar source magnitude": np.float32(s.magnitude).item(), "number of receivers": int(num_rec_tot),
        "length of seismogram  in time samples": int(nseismo),
        "seismogram sampling in sec": np.float32(np.float32(prob.deltat) * np.float32(prob.seis_it)).item(),
    }
    if dump_wavefields:
        a.update({"number of strain dumps": int(nstrain), "strain dump sampling rate in sec": float(deltat_coarse),
                  "dump type (displ_only, displ_velo, fullfields)": dump_type,
                  "kernel wavefield rmin": 0.0, "kernel wavefield rmax": spec.router / 1000.0,
                  "kernel wavefield colatmin": 0.0, "kernel wavefield colatmax": 0.0})
    else:
        a.update({"number of strain dumps": 0, "strain dump sampling rate in sec": 0.0})
    a.update({"number of snapshot dumps": 0, "snapshot dump sampling rate in sec": 0.0, "receiver components ": "cyl"})
    ib, ie, jb, je = getattr(prob, "dump_block", (0, 4, 0, 4)) if dump_type == "fullfields" else (0, 4, 0, 4)
    a.update({"ibeg": ib, "iend": ie, "jbeg": jb, "jend": je, "source shift factor in sec": np.float32(shift).item(),
              "source shift factor for deltat": int(round(shift / prob.deltat)),
              "source shift factor for seis_dt": int(round(shift / seis_dt)),
              "source shift factor for deltat_coarse": int(round(shift / deltat_coarse)) if deltat_coarse > 0 else 0,
              "receiver file type": "colatlon", "receiver spacing (0 if not even)": 0.0,
              "use netcdf for wavefield output?": "T", "percent completed": 100, "finalized": 1})
    return a


def mesh_group(probs: Sequence) -> Dict[str, np.ndarray]:
    """The Mesh group for displ_only / strain_only dumps: coordinates and material at the kwf points
    of every rank (rank blocks concatenated, each solid then fluid: nc_routines.F90:668-691,
    717-823) and, per dumped element, midpoint / corner / GLL point indices into that list
    (0-based, meshes_io.F90:641-778)."""
    from .mesh import element_coords
    from .precomp import geometry, material
    pts_s, pts_z, mat = [], [], {k: [] for k in ("vp", "vs", "rho", "lambda", "mu", "xi", "phi", "eta", "Qmu", "Qka")}
    mid, eltype, axis, fem, sem, mps, mpz = [], [], [], [], [], [], []
    base = 0
    for p in probs:
        nm = getattr(p, "nc_mesh", None)
        if nm is not None:                 # a run set up by the native pre-computation: the arrays travel in its container
            pts_s.append(nm["mesh_S"])
            pts_z.append(nm["mesh_Z"])
            for k in mat:
                mat[k].append(nm["mesh_" + k])
            mid.append(nm["midpoint_mesh"] + base)
            eltype.append(nm["eltype"])
            axis.append(nm["axis"])
            fem.append(nm["fem_mesh"].reshape(-1, 4) + base)
            sem.append(nm["sem_mesh"].reshape(-1, 5, 5) + base)
            mps.append(nm["mp_mesh_S"])
            mpz.append(nm["mp_mesh_Z"])
            base += nm["mesh_S"].size
            continue
        m, q = p.mesh, p.kwf
        npt = q["npoint_solid_kwf"] + q["npoint_fluid_kwf"]
        S, Z = np.zeros(npt), np.zeros(npt)
        M = {k: np.zeros(npt, np.float32) for k in mat}
        off = 0
        for es, nel in ((m.solid, m.nel_solid), (m.fluid, m.nel_fluid)):
            if nel == 0:
                continue
            g = geometry(es, m.basis)
            rho, lam, mu, xi, phi, eta, vp, qmu, qka = material(m.spec, es, g)
            msk = q["kwf_mask"][off:off + nel].astype(bool)
            idx = q["mapping_ijel_ikwf"][off:off + nel][msk] - 1
            S[idx] = g.s[msk]
            Z[idx] = g.z[msk]
            vals = {"rho": rho, "lambda": lam, "mu": mu, "xi": xi, "phi": phi, "eta": eta,
                    "vp": np.sqrt((lam + 2 * mu) / rho), "vs": np.sqrt(mu / rho),
                    "Qmu": np.broadcast_to(np.asarray(qmu, float)[:, None, None], rho.shape),
                    "Qka": np.broadcast_to(np.asarray(qka, float)[:, None, None], rho.shape)}
            for k, v in vals.items():
                M[k][idx] = v[msk]
            mp = q["mapping_ijel_ikwf"][off:off + nel]
            mid.append(mp[:, 2, 2] - 1 + base)
            eltype.append(np.zeros(nel, np.int32))                       # all 'curved' in the synthetic meshes
            axis.append(es.axis.astype(np.int32))
            fem.append(np.stack([mp[:, 0, 0], mp[:, 0, 4], mp[:, 4, 4], mp[:, 4, 0]], axis=1) - 1 + base)
            sem.append(mp - 1 + base)                 # sem_mesh(ipol, jpol, el) = [el][jpol][ipol] in C order
            mps.append(g.s[:, 2, 2])
            mpz.append(g.z[:, 2, 2])
            off += nel
        pts_s.append(S)
        pts_z.append(Z)
        for k in mat:
            mat[k].append(M[k])
        base += npt
    b = probs[0].mesh.basis
    out = {"mesh_S": np.concatenate(pts_s), "mesh_Z": np.concatenate(pts_z)}
    for k, v in mat.items():
        out["mesh_" + k] = np.concatenate(v).astype(np.float32)
    out.update({"midpoint_mesh": np.concatenate(mid).astype(np.int32), "eltype": np.concatenate(eltype),
                "axis": np.concatenate(axis), "fem_mesh": np.concatenate(fem).astype(np.int32),
                "sem_mesh": np.concatenate(sem).astype(np.int32),
                "mp_mesh_S": np.concatenate(mps), "mp_mesh_Z": np.concatenate(mpz),
                "G0": b.G0.astype(np.float64), "G1": b.G1.T.astype(np.float64).copy(),
                "G2": b.G2.T.astype(np.float64).copy(), "gll": b.eta.astype(np.float64), "glj": b.xi_k.astype(np.float64)})
    return out


def write_database(outdir: str, probs: Sequence, seismograms: Sequence[np.ndarray],
                   snapshots: Optional[Sequence[np.ndarray]] = None, *, colat_deg=None, names=None,
                   background_model: str = "prem_iso") -> Dict:
    """Assemble the database directory from the per-rank results of a run.
    seismograms[r]: (nseismo, num_rec_r, 3) as TimeLoop.seismograms() (or the .seis.f32 of the C++
    host); snapshots[r]: (nvars, nstrain, npoints_r) as TimeLoop.snapshots().  Receivers are
    ordered as the station list the problems were built with (Problem.rec_index)."""
    p0 = probs[0]
    dump_wavefields = snapshots is not None and p0.kwf is not None and p0.strain_it > 0
    dump_type = getattr(p0, "dump_type", "displ_only")
    mono = p0.src_order == 0
    nrec = int(sum(p.num_rec for p in probs))
    nseismo = int(max(s.shape[0] for s in seismograms))
    nstrain = int(snapshots[0].shape[1]) if dump_wavefields else 0
    npts = [int(s.shape[2]) for s in snapshots] if dump_wavefields else []
    nelem_kwf = int(sum(p.mesh.nel_solid + p.mesh.nel_fluid for p in probs))
    sch = schema(nrec=nrec, nseismo=nseismo, niter=p0.niter, dump_wavefields=dump_wavefields, dump_type=dump_type,
                 monopole=mono, npoints_global=sum(npts), nstrain=nstrain, nelem_kwf_global=nelem_kwf, anel=p0.anel)
    deltat_coarse = p0.deltat * p0.strain_it if dump_wavefields else 0.0
    sch["attributes"] = global_attributes(p0, nseismo=nseismo, nstrain=nstrain, deltat_coarse=deltat_coarse,
                                          num_rec_tot=nrec, dump_wavefields=dump_wavefields, dump_type=dump_type,
                                          background_model=background_model)
    data: Dict[str, Dict[str, np.ndarray]] = {"": {}, "Seismograms": {}, "Snapshots": {}, "Surface": {}, "Mesh": {}}
    S = data["Seismograms"]
    if nrec:
        disp = np.zeros((nrec, 3, nseismo), np.float32)
        theta = np.zeros(nrec, np.float32)
        proc = np.zeros(nrec, np.int32)
        for r, (p, s) in enumerate(zip(probs, seismograms)):
            if p.num_rec:
                disp[p.rec_index] = np.asarray(s, np.float32).transpose(1, 2, 0)
                proc[p.rec_index] = r
        if colat_deg is not None:
            theta[:] = np.asarray(colat_deg, np.float32)
        nm = names or [f"REC{k:04d}" for k in range(nrec)]
        rn = np.full((40, nrec), b" ", dtype="S1")
        for k, n in enumerate(nm):
            for c, ch in enumerate(n[:40]):
                rn[c, k] = ch.encode()
        S.update({"displacement": disp, "time": np.arange(nseismo) * p0.deltat * p0.seis_it,
                  "phi": np.zeros(nrec, np.float32), "theta_requested": theta, "theta": theta,
                  "processor_of_receiver": proc, "receiver_name": rn})
    stf = np.asarray(p0.stf, np.float32)
    dstf = np.gradient(stf, p0.deltat).astype(np.float32) if stf.size > 1 else stf
    it_seis = np.arange(nseismo) * p0.seis_it
    pad = lambda a, idx: np.where(idx >= 1, a[np.clip(idx - 1, 0, a.size - 1)], 0.0).astype(np.float32)
    S.update({"stf_seis": pad(stf, it_seis), "stf_d_seis": pad(dstf, it_seis),
              "stf_iter": stf[:p0.niter], "stf_d_iter": dstf[:p0.niter]})
    if dump_wavefields:
        it_dump = np.arange(nstrain) * p0.strain_it
        data[""]["snapshot_times"] = (it_dump * p0.deltat).astype(np.float32)
        names_v = SNAP_VARS[(dump_type, mono)]
        # the device buffer of displ_only keeps the (zero) phi plane for monopole sources
        planes = [0, 2] if (dump_type == "displ_only" and mono) else list(range(len(names_v)))
        for n, v in zip(names_v, planes):
            data["Snapshots"][n] = np.concatenate([np.asarray(s[v], np.float32) for s in snapshots], axis=1)
        data["Snapshots"].update({"stf_dump": pad(stf, it_dump), "stf_d_dump": pad(dstf, it_dump)})
        if dump_type in ("displ_only", "strain_only"):
            mg = mesh_group(probs)
            for n in sch["groups"]["Mesh"]["variables"]:
                data["Mesh"][n] = mg[n]
    os.makedirs(outdir, exist_ok=True)
    for grp, vars_ in data.items():
        defs = sch["variables"] if grp == "" else sch["groups"][grp]["variables"]
        dims_all = dict(sch["dimensions"])
        if grp:
            dims_all.update(sch["groups"][grp]["dimensions"])
            os.makedirs(os.path.join(outdir, grp), exist_ok=True)
        dims_all.update(sch["groups"]["Mesh"]["dimensions"])
        for n, arr in vars_.items():
            d = defs[n]
            shape = tuple(dims_all[k] for k in d["dims"])
            arr = np.ascontiguousarray(arr, dtype=np.dtype(d["dtype"]))
            assert arr.shape == shape, (grp, n, arr.shape, shape)
            arr.tofile(os.path.join(outdir, grp, n + ".bin"))
        missing = set(defs) - set(vars_)
        assert not missing, (grp, missing)
    with open(os.path.join(outdir, "schema.json"), "w") as f:
        json.dump(sch, f, indent=1, default=lambda o: o.item() if hasattr(o, "item") else str(o))
    return sch


def native_problem(rec: Dict[str, np.ndarray], rec_offset: int = 0):
    """What write_database needs of a rank, from the module-variable container the native pre-computation wrote
    (axisem_b200_precomp, hostcxx/precomp.cpp; read with meshdb_io.read_axbprob)."""
    from types import SimpleNamespace as NS
    from ..capi import SCHEMES, STF_TYPES
    from .source import SourceParams
    i = lambda k, d=None: int(np.asarray(rec[k]).reshape(-1)[0]) if k in rec else d
    f = lambda k: float(np.asarray(rec[k]).reshape(-1)[0])
    text = lambda k: "".join(chr(int(c)) for c in np.asarray(rec[k]).reshape(-1))
    order = i("data_source%src_order")
    src = SourceParams(src_type2=text("data_source%src_type2"), depth=f("data_source%src_depth"), magnitude=f("data_source%magnitude"),
                       stf_type=text("data_source%stf_name"), t_0=f("data_source%t_0"), decay=f("data_source%decay"),
                       shift_seconds=f("data_source%shift_fact"))
    assert STF_TYPES[src.stf_type] == i("data_source%stf_type")
    basis = NS(G0=np.asarray(rec["data_spec%G0"], np.float64).reshape(-1),
               G1=np.asarray(rec["data_spec%G1"], np.float64).reshape(5, 5).T, G2=np.asarray(rec["data_spec%G2"], np.float64).reshape(5, 5).T,
               eta=np.asarray(rec["data_spec%eta"], np.float64).reshape(-1), xi_k=np.asarray(rec["data_spec%xi_k"], np.float64).reshape(-1))
    mesh = NS(nel_solid=i("data_mesh%nel_solid"), nel_fluid=i("data_mesh%nel_fluid"), basis=basis,
              spec=NS(router=f("data_mesh%router")))
    dump = i("data_io%dump_wavefields", 0) != 0
    kwf = dict(npoint_solid_kwf=i("data_mesh%npoint_solid_kwf"), npoint_fluid_kwf=i("data_mesh%npoint_fluid_kwf")) if dump else None
    num_rec = i("data_mesh%num_rec")
    scheme = {v: k for k, v in SCHEMES.items()}[i("data_time%time_scheme")]
    p = NS(mesh=mesh, source=src, src_order=order, src_type=("monopole", "dipole", "quadpole")[order], time_scheme=scheme,
           deltat=f("data_time%deltat"), niter=i("data_time%niter"), seis_it=i("data_time%seis_it"), strain_it=i("data_time%strain_it"),
           stf=np.asarray(rec["data_source%stf"], np.float32).reshape(-1), anel=i("attenuation%anel_true", 0) != 0,
           num_rec=num_rec, rec_index=(np.asarray(rec["data_mesh%loc2globrec"]).reshape(-1) - 1 if num_rec else np.zeros(0, int)),
           kwf=kwf, dump_type="displ_only")
    if dump:
        p.nc_mesh = {k.split("%", 1)[1]: np.asarray(v) for k, v in rec.items() if k.startswith("nc_mesh%")}
    return p


def write_database_native(outdir: str, containers: Sequence[str], run_prefix: str, *, colat_deg=None, names=None,
                          background_model: str = "prem_iso") -> Dict:
    """The database directory of a run of the native chain: `containers` are the PREFIX.rankNNNN.axbp files
    of axisem_b200_precomp, `run_prefix` the --out of axisem_b200_solver (RUN.rankNNNN.seis.f32 / .snap.f32)."""
    from .meshdb_io import read_axbprob
    probs = [native_problem(read_axbprob(c)) for c in containers]
    seis, snaps = [], []
    for r, p in enumerate(probs):
        fs = f"{run_prefix}.rank{r:04d}.seis.f32"
        seis.append(np.fromfile(fs, dtype=np.float32).reshape(-1, p.num_rec, 3) if p.num_rec else np.zeros((0, 0, 3), np.float32))
        if p.kwf is not None:
            npt = p.kwf["npoint_solid_kwf"] + p.kwf["npoint_fluid_kwf"]
            snaps.append(np.fromfile(f"{run_prefix}.rank{r:04d}.snap.f32", dtype=np.float32).reshape(3, -1, npt))
    if colat_deg is None:
        colat_deg = np.asarray(read_axbprob(containers[0])["data_mesh%recfile_readth"], np.float64).reshape(-1)
    return write_database(outdir, probs, seis, snaps or None, colat_deg=colat_deg, names=names, background_model=background_model)


def read_variable(outdir: str, group: str, name: str) -> np.ndarray:
    sch = json.load(open(os.path.join(outdir, "schema.json")))
    d = (sch["variables"] if not group else sch["groups"][group]["variables"])[name]
    dims = dict(sch["dimensions"])
    for g in sch["groups"].values():
        dims.update(g["dimensions"])
    shape = tuple(dims[k] for k in d["dims"])
    return np.fromfile(os.path.join(outdir, group, name + ".bin"), dtype=np.dtype(d["dtype"])).reshape(shape)
