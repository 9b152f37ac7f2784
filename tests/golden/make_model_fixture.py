#!/usr/bin/env python
"""Turns the reference's own tabulated background model of TEST04
(/root/reference/TESTING/TEST04_anelastic_anisotropic/model.bm: prem_ani sampled by the
reference, columns radius rho vpv vsv vph vsh eta qka qmu) into tests/golden/prem_ani_model_bm.npz.
Run in the build container, where /root/reference exists; the .npz travels."""
import os
import numpy as np

SRC = "/root/reference/TESTING/TEST04_anelastic_anisotropic/model.bm"
rows = []
for line in open(SRC):
    t = line.split()
    if len(t) == 9:
        try:
            rows.append([float(v) for v in t])
        except ValueError:
            pass
a = np.array(rows)
SRC_ISO = "/root/reference/TESTING/TEST01_elastic_isotropic/model.bm"
rows_iso = []
for line in open(SRC_ISO):
    t = line.split()
    if len(t) == 4:
        try:
            rows_iso.append([float(v) for v in t])
        except ValueError:
            pass
b = np.array(rows_iso)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prem_ani_model_bm.npz")
np.savez_compressed(out, table=a, table_iso=b, columns=np.array(["radius", "rho", "vpv", "vsv", "vph", "vsh", "eta", "qka", "qmu"]))
print(out, a.shape)
