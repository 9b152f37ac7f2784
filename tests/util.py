"""Shared helpers for the parity tests."""
import numpy as np

from axisem_b200.host import (AttenuationModel, SourceParams, build_problem,
                              prem_mesh_spec)


def small_spec(ntheta=16, nr=18, anisotropic=False):
    return prem_mesh_spec(ntheta=ntheta, nr_target=nr, anisotropic=anisotropic)


def make_problem(src="explosion", anel=False, ntheta=16, nr=18, niter=40, rank=0, nranks=1,
                 scheme="newmark2", anisotropic=False, dump=False, strain_it=0, t_0=40.0,
                 seis_it=1, coarse_grained=True, nranks_r=1, source_kw=None):
    spec = small_spec(ntheta, nr, anisotropic)
    att = AttenuationModel(coarse_grained=coarse_grained) if anel else None
    return build_problem(spec, SourceParams(src_type2=src, t_0=t_0, **(source_kw or {})), anel=anel, att=att, niter=niter,
                         rank=rank, nranks=nranks, time_scheme=scheme, dump=dump,
                         strain_it=strain_it, seis_it=seis_it, nranks_r=nranks_r)


def seeded_state(loop, seed=1234, scale=1e-3, fields=("disp", "velo", "acc0", "chi", "dchi", "ddchi0")):
    """N(0,1)*scale initial fields, identical for every backend."""
    rng = np.random.default_rng(seed)
    out = {}
    for f in fields:
        shp = loop._field_shape(f)
        if 0 in shp:
            continue
        out[f] = (rng.standard_normal(shp) * scale).astype(np.float32)
    return out


def apply_state(loop, state):
    for k, v in state.items():
        loop.set(k, v)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.sqrt(np.sum((a - b) ** 2))
    n = np.sqrt(np.sum(b ** 2))
    return d / n if n > 0 else d
