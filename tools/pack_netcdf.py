#!/usr/bin/env python
"""Turns a database directory written by axisem_b200.host.nc_layout.write_database (schema.json +
raw arrays) into the NetCDF-4 file the reference's readers expect (axisem_output.nc4).

    python tools/pack_netcdf.py RUN.ncdir axisem_output.nc4

Needs netCDF4 or h5py (neither is in the build image; any workstation with Instaseis has them).
With h5py the dimensions become HDF5 dimension scales, which is what libnetcdf writes itself."""
import json
import os
import sys

import numpy as np


def load(outdir, group, name, d, dims):
    shape = tuple(dims[k] for k in d["dims"])
    return np.fromfile(os.path.join(outdir, group, name + ".bin"), dtype=np.dtype(d["dtype"])).reshape(shape)


def main():
    outdir, ncfile = sys.argv[1:3]
    sch = json.load(open(os.path.join(outdir, "schema.json")))
    dims = dict(sch["dimensions"])
    for g in sch["groups"].values():
        dims.update(g["dimensions"])
    try:
        import netCDF4
    except ImportError:
        netCDF4 = None
    if netCDF4 is not None:
        with netCDF4.Dataset(ncfile, "w", format="NETCDF4") as nc:
            for k, v in sch["attributes"].items():
                nc.setncattr(k, v)
            for k, n in sch["dimensions"].items():
                nc.createDimension(k, n)
            handles = {"": nc}
            for gname, g in sch["groups"].items():
                h = nc.createGroup(gname)
                handles[gname] = h
                for k, n in g["dimensions"].items():
                    h.createDimension(k, n)
            for gname, vars_ in [("", sch["variables"])] + [(k, g["variables"]) for k, g in sch["groups"].items()]:
                for name, d in vars_.items():
                    kw = {}
                    if "chunks" in d:
                        kw["chunksizes"] = d["chunks"]
                    fill = d.get("attrs", {}).get("_FillValue")
                    v = handles[gname].createVariable(name, "S1" if d["dtype"] == "S1" else d["dtype"], d["dims"],
                                                      fill_value=fill, **kw)
                    for ak, av in d.get("attrs", {}).items():
                        if ak != "_FillValue":
                            v.setncattr(ak, av)
                    v[...] = load(outdir, gname, name, d, dims)
        return
    import h5py                                    # raises if neither library is present
    with h5py.File(ncfile, "w") as f:
        for k, v in sch["attributes"].items():
            f.attrs[k] = v
        scales = {}
        for k, n in sch["dimensions"].items():
            scales[k] = f.create_dataset(k, shape=(n,), dtype="f4")
            scales[k].make_scale(k)
        for gname, g in sch["groups"].items():
            h = f.create_group(gname)
            for k, n in g["dimensions"].items():
                scales[k] = h.create_dataset(k, shape=(n,), dtype="f4")
                scales[k].make_scale(k)
        for gname, vars_ in [("", sch["variables"])] + [(k, g["variables"]) for k, g in sch["groups"].items()]:
            h = f if not gname else f[gname]
            for name, d in vars_.items():
                ds = h.create_dataset(name, data=load(outdir, gname, name, d, dims),
                                      chunks=tuple(d["chunks"]) if "chunks" in d else None)
                for k, dn in enumerate(d["dims"]):
                    ds.dims[k].attach_scale(scales[dn])
                for ak, av in d.get("attrs", {}).items():
                    ds.attrs[ak] = av


if __name__ == "__main__":
    main()
