#!/usr/bin/env python
"""Output database (the layout of axisem_output.nc4, axisem_b200/host/nc_layout.py) of a run of the native
chain:

    axisem_b200_precomp --out PRE [...] --strain-it K meshdb.dat0000 ...
    axisem_b200_solver  --out RUN PRE.rank0000.axbp ...
    tools/native_to_nc_layout.py --out DB.ncdir --run RUN [--model prem_ani] PRE.rank0000.axbp ...
    tools/pack_netcdf.py DB.ncdir axisem_output.nc4        (where netCDF4 or h5py exists)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--run", required=True)
    ap.add_argument("--model", default="prem_iso")
    ap.add_argument("containers", nargs="+")
    a = ap.parse_args()
    from axisem_b200.host import nc_layout
    sch = nc_layout.write_database_native(a.out, a.containers, a.run, background_model=a.model)
    nvar = len(sch["variables"]) + sum(len(g["variables"]) for g in sch["groups"].values())
    print(f"{a.out}: {nvar} variables, {len(sch['attributes'])} global attributes")


if __name__ == "__main__":
    main()
