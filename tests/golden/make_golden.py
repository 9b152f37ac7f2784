#!/usr/bin/env python
"""Regenerates the committed golden vectors tests/golden/case_*.npz.

Each file freezes the complete C-ABI inputs of one tiny configuration (synthetic PREM-type
mesh, 4 theta columns) together with what the CPU oracle (oracle/axisem_oracle.c, the
restatement of the reference time loop) produced for them: seismograms, wavefield
snapshots and the final state.  tests/test_golden.py replays the inputs through the
oracle (CPU, bit-exact) and tests/test_gpu_golden.py through the CUDA library.

The reference itself cannot be run here (Fortran; no compiler in the image) and ships no
array-level vectors for this path, so these vectors pin the *oracle*, not the Fortran —
see DESIGN.md section 2 ("parity unpinned").

    python tests/golden/make_golden.py [case names]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from axisem_b200.host import (AttenuationModel, SourceParams, SpectralBasis, build_problem, prem_mesh_spec,  # noqa: E402
                              stable_timestep)
from axisem_b200.host.problem_io import save_problem  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {
    # name: (source, anel, anisotropic, scheme, nsteps)
    "mono_elastic_iso": ("explosion", False, False, "newmark2", 40),
    "dipole_anel_iso": ("mtr", True, False, "newmark2", 40),
    "quad_anel_ani": ("mtp", True, True, "newmark2", 40),
    "dipole_elastic_symplec4": ("thetaforce", False, True, "symplec4", 12),
    # COARSE_GRAINED false: memory variables at all 25 points (attenuation.f90:210-334)
    "dipole_anel_full_ani": ("mtr", "full", True, "newmark2", 40),
}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])
    for name, (src, anel, ani, scheme, n) in CASES.items():
        if only and name not in only:
            continue
        att = AttenuationModel(coarse_grained=False) if anel == "full" else None
        anel = bool(anel)
        spec = prem_mesh_spec(ntheta=4, nr_target=9, anisotropic=ani)
        dt = stable_timestep(spec, SpectralBasis(4)) * (1.5 if scheme != "newmark2" else 1.0)
        # a short source so that the STF peaks inside the run
        prob = build_problem(spec, SourceParams(src_type2=src, t_0=8.0 * dt), anel=anel, att=att, niter=n,
                             time_scheme=scheme, dump=True, strain_it=8, seis_it=2,
                             rec_colat_deg=[10.0, 60.0, 120.0, 170.0], threads=1)
        loop = oracle.make_loop(prob)
        # a seeded non-zero start so every term is exercised from step 1
        rng = np.random.default_rng(7)
        init = {}
        for f in ("disp", "velo", "chi", "dchi"):
            shp = loop._field_shape(f)
            v = (rng.standard_normal(shp) * 1e-3).astype(np.float32)
            if f in ("disp", "velo") and src == "explosion":
                v[1] = 0.0
            loop.set(f, v)
            init["init_" + f] = v
        loop.run(n)
        extra = dict(init)
        extra["seismograms"] = loop.seismograms()
        extra["snapshots"] = loop.snapshots()
        for f in ("disp", "velo", "chi", "dchi") + (("memvar",) if anel else ()):
            extra["final_" + f] = loop.get(f)
        extra["nsteps"] = np.int32(n)
        path = os.path.join(here, f"case_{name}.npz")
        save_problem(prob, path, extra)
        print(name, prob.mesh.nel_solid, prob.mesh.nel_fluid, os.path.getsize(path) // 1024, "KiB",
              "dt = %.3f s, max|stf| = %.3e, max|seis| = %.3e"
              % (prob.deltat, np.abs(prob.stf).max(), np.abs(extra["seismograms"]).max()))


if __name__ == "__main__":
    main()
