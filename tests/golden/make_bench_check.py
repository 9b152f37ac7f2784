"""Golden of bench.py's correctness check: the mid-size dipole + cg4-attenuation case run by the
CPU oracle on ONE rank; bench.py runs the same case on its N ranks through the product library
and reports the relative L2 difference of the seismograms (`check` in the JSON line).

    python tests/golden/make_bench_check.py        (writes tests/golden/bench_check_mtr_cg4.npz)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from axisem_b200.host import SourceParams, build_problem, prem_mesh_spec  # noqa: E402
from oracle import oracle  # noqa: E402

NTHETA, NR, T_0, NITER, SEIS_IT, R_MIN_KM = 64, 48, 25.0, 9600, 24, 800.0
COLAT = np.linspace(4.0, 176.0, 24)


def main():
    spec = prem_mesh_spec(ntheta=NTHETA, nr_target=NR, r_min_km=R_MIN_KM)
    prob = build_problem(spec, SourceParams(src_type2="mtr", t_0=T_0), anel=True, niter=NITER,
                         rec_colat_deg=COLAT, seis_it=SEIS_IT)
    print("dt", prob.deltat, "elements", spec.nelem, "stations", prob.num_rec)
    O = oracle.make_loop(prob)
    O.run(NITER)
    s = O.seismograms()
    peak = np.abs(s).max(axis=(0, 2))
    print("peak per station", peak)
    assert (peak > 1e-4 * peak.max()).all(), "every station must carry signal"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bench_check_mtr_cg4.npz"),
                        seismograms=s.astype(np.float32), ntheta=NTHETA, nr=NR, t_0=T_0, niter=NITER,
                        seis_it=SEIS_IT, colat_deg=COLAT, deltat=prob.deltat, r_min_km=R_MIN_KM)


if __name__ == "__main__":
    main()
