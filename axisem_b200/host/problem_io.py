"""Freeze / thaw one rank's time-loop inputs (a `Problem`) to a single .npz file.

What is stored is exactly what crosses the C ABI of include/axisem_b200.h — the arrays
`prepare_waves` (SOLVER/time_evol_wave.F90:47-225) leaves in the Fortran modules — so a
frozen problem can be replayed through any implementation of the header without the
mesher / pre-computation code.  Used for the committed golden vectors (tests/golden/).
"""
from __future__ import annotations

import json
from types import SimpleNamespace

import numpy as np

_MESH_ARRAYS = ["igloc_solid", "igloc_fluid", "axis_solid", "axis_fluid", "ax_el_solid",
                "ax_el_fluid", "bdry_solid_el", "bdry_fluid_el", "bdry_jpol_solid",
                "bdry_jpol_fluid"]
_MESH_SCALARS = ["rank", "nranks", "nel_solid", "nel_fluid", "nglob_solid", "nglob_fluid", "nel_bdry"]
_HALO_ARRAYS = ["list_peer", "sizemsg", "glocal_index_msg", "glob2el"]
_PROB_ARRAYS = ["inv_mass_rho", "inv_mass_fluid", "fluid_free_surface_mask", "inv_rho_fluid",
                "bdry_matr", "ielsrc", "source_term_el", "stf", "recfile_el", "rec_index",
                "solid_absorbing_gamma", "fluid_absorbing_gamma"]
_PROB_SCALARS = ["src_type", "src_order", "time_scheme", "deltat", "niter", "seis_it", "strain_it",
                 "anel", "nelsrc"]
_SRC_FIELDS = ["src_type2", "depth", "magnitude", "stf_type", "t_0", "decay", "shift_fact", "shift_seconds", "discrete_choice"]
_DICTS = ["solid", "fluid", "pw_solid", "pw_fluid", "att", "kwf"]


def save_problem(prob, path: str, extra=None):
    """Write `prob` (+ optional dict of expected outputs `extra`) to `path` (.npz)."""
    arrs, meta = {}, {"mesh": {}, "prob": {}, "source": {}, "dict_scalars": {}}
    m = prob.mesh
    for k in _MESH_ARRAYS:
        arrs["mesh/" + k] = np.asarray(getattr(m, k))
    for k in _MESH_SCALARS:
        meta["mesh"][k] = int(getattr(m, k))
    for g in ("G0", "G1", "G1T", "G2", "G2T"):
        arrs["basis/" + g] = np.asarray(getattr(m.basis, g))
    for side in ("halo_solid", "halo_fluid"):
        h = getattr(m, side)
        meta["mesh"][side] = {"nmsg": int(h.nmsg), "num_comm_gll": int(h.num_comm_gll)}
        for k in _HALO_ARRAYS:
            arrs[f"{side}/{k}"] = np.asarray(getattr(h, k))
    for k in _PROB_ARRAYS:
        v = getattr(prob, k)
        if v is not None:
            arrs["prob/" + k] = np.asarray(v)
    for k in _PROB_SCALARS:
        v = getattr(prob, k)
        meta["prob"][k] = v if isinstance(v, (str, bool)) else (float(v) if isinstance(v, float) else int(v))
    for k in _SRC_FIELDS:
        meta["source"][k] = getattr(prob.source, k)
    for d in _DICTS:
        dd = getattr(prob, d)
        if dd is None:
            continue
        meta["dict_scalars"][d] = {}
        for k, v in dd.items():
            if isinstance(v, np.ndarray):
                arrs[f"{d}/{k}"] = v
            else:
                meta["dict_scalars"][d][k] = (bool(v) if isinstance(v, (bool, np.bool_)) else
                                              int(v) if isinstance(v, (int, np.integer)) else float(v))
    for k, v in (extra or {}).items():
        arrs["expect/" + k] = np.asarray(v)
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrs)


class FrozenProblem(SimpleNamespace):
    @property
    def num_rec(self):
        return int(self.recfile_el.shape[0])


def load_problem(path: str):
    """-> (problem usable by axisem_b200.capi.TimeLoop, dict of expected outputs)."""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    mesh = SimpleNamespace(**{k: meta["mesh"][k] for k in _MESH_SCALARS})
    for k in _MESH_ARRAYS:
        setattr(mesh, k, z["mesh/" + k])
    mesh.basis = SimpleNamespace(**{g: z["basis/" + g] for g in ("G0", "G1", "G1T", "G2", "G2T")})
    for side in ("halo_solid", "halo_fluid"):
        h = SimpleNamespace(**meta["mesh"][side])
        for k in _HALO_ARRAYS:
            setattr(h, k, z[f"{side}/{k}"])
        setattr(mesh, side, h)
    p = FrozenProblem(mesh=mesh, **meta["prob"])
    for k in _PROB_ARRAYS:
        setattr(p, k, z["prob/" + k] if "prob/" + k in z.files else None)
    p.source = SimpleNamespace(**meta["source"])
    for d in _DICTS:
        if d not in meta["dict_scalars"]:
            setattr(p, d, None)
            continue
        dd = dict(meta["dict_scalars"][d])
        for f in z.files:
            if f.startswith(d + "/"):
                dd[f[len(d) + 1:]] = z[f]
        setattr(p, d, dd)
    expect = {f[len("expect/"):]: z[f] for f in z.files if f.startswith("expect/")}
    return p, expect
