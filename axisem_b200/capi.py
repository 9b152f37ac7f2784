"""ctypes binding of the C ABI declared in include/axisem_b200.h.

The same binding drives two implementations of the header:

  * ``libaxisem_b200.so``  (prefix ``axb_``) — the CUDA sm_100a product, loaded by
    :mod:`axisem_b200.solver`;
  * the CPU test oracle, a second implementation of the same header under another symbol
    prefix; it lives outside this package and is loaded only by the tests / smoke /
    bench baseline through its own loader, never from here.

`TimeLoop` is the host-side mirror of the reference's `time_loop` seam
(SOLVER/time_evol_wave.F90:231): hand over the module arrays once, `run`, fetch the
seismogram / snapshot buffers.  It takes a :class:`axisem_b200.host.problem.Problem`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

_F = C.POINTER(C.c_float)
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)

SOLID_FIELDS = ["M11s", "M21s", "M41s", "M12s", "M22s", "M32s", "M42s", "M11z", "M21z", "M41z",
                "M13s", "M33s", "M43s", "M1phi", "M2phi", "M4phi",
                "M_1", "M_2", "M_3", "M_4", "M_5", "M_6", "M_7", "M_8",
                "M_w1", "M_w2", "M_w3", "M_w4", "M_w5",
                "M0_w1", "M0_w2", "M0_w3", "M0_w4", "M0_w5", "M0_w6", "M0_w7", "M0_w8",
                "M0_w9", "M0_w10"]


class SolidTerms(C.Structure):
    _fields_ = [(n, _F) for n in SOLID_FIELDS]


ATT_F = ["Q_mu", "Q_kappa", "delta_mu_cg4", "delta_kappa_cg4", "Y_cg4", "V_s_eta_cg4",
         "V_s_xi_cg4", "V_z_eta_cg4", "V_z_xi_cg4", "DsDeta_over_J_sol_cg4",
         "DzDeta_over_J_sol_cg4", "DsDxi_over_J_sol_cg4", "DzDxi_over_J_sol_cg4",
         "delta_mu", "delta_kappa", "Y", "V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi",
         "Y0", "V0_s_eta", "V0_s_xi", "V0_z_eta", "V0_z_xi",
         "DsDeta_over_J_sol", "DzDeta_over_J_sol", "DsDxi_over_J_sol", "DzDxi_over_J_sol",
         "inv_s_solid"]


class Attenuation(C.Structure):
    _fields_ = ([("coarse_grained", C.c_int32), ("n_sls", C.c_int32), ("do_corr_lowq", C.c_int32),
                 ("y_j", _D), ("exp_w_j_deltat", _D), ("ts_fac_t", _D), ("ts_fac_tm1", _D)]
                + [(n, _F) for n in ATT_F])


SCHEMES = {"newmark2": 0, "symplec4": 1, "ML_SO4m5": 2, "ML_SO6m7": 3, "KL_O8m17": 4,
           "SS_35o10": 5}
# sub-stages per step (symplectic_coefficients, time_evol_wave.F90:749-967)
NSTAGES = {"symplec4": 4, "ML_SO4m5": 5, "ML_SO6m7": 7, "KL_O8m17": 17, "SS_35o10": 35}
# stf_type codes of include/axisem_b200.h (dirac_1 is a Newmark-only alias of dirac_0: compute_stf_t has no case for it)
STF_TYPES = {"gauss_0": 0, "gauss_1": 1, "gauss_2": 2, "errorf": 3, "dirac_0": 4, "quheavi": 5, "dirac_1": 4}
FIELDS = {"disp": 0, "velo": 1, "acc0": 2, "acc1": 3, "chi": 4, "dchi": 5, "ddchi0": 6,
          "ddchi1": 7, "memvar": 8, "src_dev_tm1": 9, "src_tr_tm1": 10}
DUMP_TYPES = {"displ_only": 0, "strain_only": 1, "fullfields": 2}
OPS = {"solid_stiffness": 0, "anel_stiffness": 1, "fluid_stiffness": 2, "pdistsum_solid": 3,
       "pdistsum_fluid": 4, "memvars": 5, "bdry2fluid": 6, "bdry2solid": 7}

# every symbol include/axisem_b200.h declares
SYMBOLS = ["last_error", "create", "destroy", "set_mesh", "set_solid_terms", "set_fluid_terms",
           "set_mass", "set_energy", "set_sponge", "set_sf_boundary", "set_attenuation", "set_source",
           "set_stf_values", "set_stf_params", "get_stf_symp", "set_receivers", "set_kwf", "set_dump", "snapshot_layout", "set_xdmf", "xdmf_count", "fetch_xdmf", "set_halo", "set_time",
           "finalize_setup", "set_stream", "synchronize", "connect_local", "ipc_blob_bytes", "ipc_export", "ipc_import", "run", "run_group",
           "profile", "get_profile", "iter", "nseismo", "nstrain", "gpu_launches", "fetch_seismograms",
           "fetch_snapshots", "fetch_energy", "get_state", "set_state", "apply_op"]


class AxbError(RuntimeError):
    pass


def _fp(a: Optional[np.ndarray], keep: list):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    keep.append(a)
    return a.ctypes.data_as(_F)


def _ip(a: Optional[np.ndarray], keep: list):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.int32)
    keep.append(a)
    return a.ctypes.data_as(_I)


def _dp(a: Optional[np.ndarray], keep: list):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    keep.append(a)
    return a.ctypes.data_as(_D)


def fortran_matrix(G: np.ndarray) -> np.ndarray:
    """numpy G[a,b] (== Fortran G(a,b)) -> flat Fortran memory order (a fastest)."""
    return np.ascontiguousarray(np.asarray(G, dtype=np.float32).T).reshape(-1)


class Library:
    """A loaded implementation of the header (product or oracle)."""

    def __init__(self, path: str, prefix: str):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        missing = [s for s in SYMBOLS if not hasattr(self.lib, prefix + s)]
        if missing:
            raise AxbError(f"{path}: missing symbols {missing}")
        self.fn = {s: getattr(self.lib, prefix + s) for s in SYMBOLS}
        self.fn["last_error"].restype = C.c_char_p
        self.fn["gpu_launches"].restype = C.c_int64
        for s in ("iter", "nseismo", "nstrain"):
            self.fn[s].restype = C.c_int32

    def check(self, rc: int):
        if rc != 0:
            raise AxbError(self.fn["last_error"]().decode())


class TimeLoop:
    """One rank's device-resident (or oracle) time loop."""

    def __init__(self, lib: Library, prob, device: int = 0):
        self.lib = lib
        self.prob = prob
        self._keep: list = []
        self.h = C.c_void_p()
        m = prob.mesh
        lib.check(lib.fn["create"](C.byref(self.h), C.c_int32(device), C.c_int32(m.rank),
                                   C.c_int32(m.nranks)))
        self._setup()

    # ------------------------------------------------------------------------------
    def _setup(self):
        p, lib, k, h = self.prob, self.lib, self._keep, self.h
        m = p.mesh
        b = m.basis
        ck = lib.check
        fn = lib.fn
        ck(fn["set_mesh"](h, 4, m.nel_solid, m.nel_fluid, m.nglob_solid, m.nglob_fluid,
                          _ip(m.igloc_solid, k), _ip(m.igloc_fluid, k),
                          _ip(m.axis_solid, k), _ip(m.axis_fluid, k),
                          _ip(m.ax_el_solid, k), C.c_int32(m.ax_el_solid.size),
                          _ip(m.ax_el_fluid, k), C.c_int32(m.ax_el_fluid.size),
                          _fp(b.G0, k), _fp(fortran_matrix(b.G1), k), _fp(fortran_matrix(b.G1T), k),
                          _fp(fortran_matrix(b.G2), k), _fp(fortran_matrix(b.G2T), k)))
        st = SolidTerms()
        for n in SOLID_FIELDS:
            setattr(st, n, _fp(p.solid.get(n), k))
        ck(fn["set_solid_terms"](h, C.c_int32(p.src_order), C.byref(st)))
        if m.nel_fluid:
            f = p.fluid
            ck(fn["set_fluid_terms"](h, _fp(f["M1chi_fl"], k), _fp(f["M2chi_fl"], k),
                                     _fp(f["M4chi_fl"], k), _fp(f.get("M_w_fl"), k),
                                     _fp(f.get("M0_w_fl"), k), _fp(p.inv_mass_fluid, k),
                                     _fp(p.fluid_free_surface_mask, k)))
        ck(fn["set_mass"](h, _fp(p.inv_mass_rho, k)))
        if getattr(p, "unassem_mass_rho_solid", None) is not None:
            ck(fn["set_energy"](h, _fp(p.unassem_mass_rho_solid, k), _fp(p.unassem_mass_lam_fluid, k)))
        if p.solid_absorbing_gamma is not None or p.fluid_absorbing_gamma is not None:
            ck(fn["set_sponge"](h, _fp(p.solid_absorbing_gamma, k), _fp(p.fluid_absorbing_gamma, k)))
        if m.nel_bdry:
            ck(fn["set_sf_boundary"](h, C.c_int32(m.nel_bdry), _ip(m.bdry_solid_el, k),
                                     _ip(m.bdry_fluid_el, k), _ip(m.bdry_jpol_solid, k),
                                     _ip(m.bdry_jpol_fluid, k), _fp(p.bdry_matr, k)))
        if p.anel:
            a = Attenuation()
            d = p.att
            a.coarse_grained = int(d["coarse_grained"])
            a.n_sls = int(d["n_sls"])
            a.do_corr_lowq = int(d["do_corr_lowq"])
            a.y_j = _dp(d["y_j"], k)
            a.exp_w_j_deltat = _dp(d["exp_w_j_deltat"], k)
            a.ts_fac_t = _dp(d["ts_fac_t"], k)
            a.ts_fac_tm1 = _dp(d["ts_fac_tm1"], k)
            src = {"Q_mu": d["Q_mu"], "Q_kappa": d["Q_kappa"], "inv_s_solid": p.pw_solid["inv_s"]}
            if a.coarse_grained:
                for n in ("delta_mu_cg4", "delta_kappa_cg4"):
                    src[n] = d[n]
                for n in ("Y_cg4", "V_s_eta_cg4", "V_s_xi_cg4", "V_z_eta_cg4", "V_z_xi_cg4"):
                    src[n] = p.solid[n]
                for n in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
                    src[n + "_over_J_sol_cg4"] = d[n + "_over_J_cg4"]
            else:
                for n in ("delta_mu", "delta_kappa"):
                    src[n] = d[n]
                for n in ("Y", "V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi",
                          "Y0", "V0_s_eta", "V0_s_xi", "V0_z_eta", "V0_z_xi"):
                    src[n] = p.solid[n]
                for n in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
                    src[n + "_over_J_sol"] = p.pw_solid[n + "_over_J"]
            for n in ATT_F:
                setattr(a, n, _fp(src.get(n), k))
            ck(fn["set_attenuation"](h, C.byref(a)))
        ck(fn["set_source"](h, C.c_int32(int(getattr(p, "fluid_src", False))), C.c_int32(p.nelsrc), _ip(p.ielsrc, k),
                            _fp(p.source_term_el, k), _fp(p.stf, k), C.c_int32(p.stf.size)))
        s = p.source
        from .host.source import stf_shift
        shift = stf_shift(s, p.deltat)
        ck(fn["set_stf_params"](h, C.c_int32(STF_TYPES[s.stf_type]), C.c_double(s.decay),
                                C.c_double(s.t_0), C.c_double(shift), C.c_double(s.magnitude)))
        # recfile_el(num_rec,3) Fortran order
        rec = np.ascontiguousarray(p.recfile_el.T)
        ck(fn["set_receivers"](h, C.c_int32(p.num_rec), _ip(rec, k)))
        if p.kwf is not None and p.strain_it > 0:
            q = p.kwf
            pf = p.pw_fluid
            ck(fn["set_kwf"](h, _ip(q["kwf_mask"], k), _ip(q["mapping_ijel_ikwf"], k),
                             C.c_int32(q["npoint_solid_kwf"]), C.c_int32(q["npoint_fluid_kwf"]),
                             _fp(p.inv_rho_fluid, k), _fp(pf["DsDeta_over_J"], k),
                             _fp(pf["DzDeta_over_J"], k), _fp(pf["DsDxi_over_J"], k),
                             _fp(pf["DzDxi_over_J"], k)))
        dump_type = getattr(p, "dump_type", "displ_only")
        if dump_type != "displ_only":
            ib, ie, jb, je = getattr(p, "dump_block", (0, 4, 0, 4))
            ps = p.pw_solid
            ck(fn["set_dump"](h, C.c_int32(DUMP_TYPES[dump_type]), C.c_int32(ib), C.c_int32(ie), C.c_int32(jb),
                              C.c_int32(je), _fp(ps["DsDeta_over_J"], k), _fp(ps["DzDeta_over_J"], k),
                              _fp(ps["DsDxi_over_J"], k), _fp(ps["DzDxi_over_J"], k), _fp(ps["inv_s"], k),
                              _fp(p.pw_fluid.get("inv_s"), k)))
        x = getattr(p, "xdmf", None)
        if x is not None:
            ps, pf = p.pw_solid, p.pw_fluid
            ck(fn["set_xdmf"](h, C.c_int32(x["snap_it"]), C.c_int32(x["i_arr"].size), C.c_int32(x["j_arr"].size),
                              _ip(x["i_arr"], k), _ip(x["j_arr"], k), _ip(x["plotting_mask"], k),
                              _ip(x["mapping_ijel_iplot"], k), C.c_int32(x["npoint_plot"]),
                              _fp(ps["DsDeta_over_J"], k), _fp(ps["DzDeta_over_J"], k), _fp(ps["DsDxi_over_J"], k),
                              _fp(ps["DzDxi_over_J"], k), _fp(ps["inv_s"], k),
                              _fp(pf.get("DsDeta_over_J"), k), _fp(pf.get("DzDeta_over_J"), k),
                              _fp(pf.get("DsDxi_over_J"), k), _fp(pf.get("DzDxi_over_J"), k),
                              _fp(pf.get("inv_s"), k), _fp(p.inv_rho_fluid, k)))
        for dom, hs in ((0, m.halo_solid), (1, m.halo_fluid)):
            if hs.nmsg:
                maxmsg = hs.glocal_index_msg.shape[1]
                g2e = np.ascontiguousarray(hs.glob2el.T)          # (3, ncomm) == Fortran (ncomm,3)
                ck(fn["set_halo"](h, C.c_int32(dom), C.c_int32(hs.nmsg), _ip(hs.list_peer, k),
                                  _ip(hs.sizemsg, k), _ip(hs.glocal_index_msg, k),
                                  C.c_int32(maxmsg), C.c_int32(hs.num_comm_gll), _ip(g2e, k)))
        ck(fn["set_time"](h, C.c_int32(SCHEMES[p.time_scheme]), C.c_double(p.deltat),
                          C.c_int32(p.niter), C.c_int32(p.seis_it), C.c_int32(p.strain_it)))
        ck(fn["finalize_setup"](h))
        self._keep.clear()          # arrays were copied by the library

    # ------------------------------------------------------------------------------
    def run(self, nsteps: int, sync: bool = True):
        """Advance nsteps; kernels are enqueued asynchronously, `sync` waits for them."""
        self.lib.check(self.lib.fn["run"](self.h, C.c_int32(nsteps)))
        if sync:
            self.synchronize()

    def synchronize(self):
        self.lib.check(self.lib.fn["synchronize"](self.h))

    def profile(self, enable=True):
        """0/False off, 1/True events around every launch, 2 around the solid element kernel only."""
        self.lib.check(self.lib.fn["profile"](self.h, C.c_int32(int(enable))))

    def get_profile(self):
        """(ms[8], launches[8]) per kernel class, see include/axisem_b200.h."""
        ms = (C.c_double * 8)()
        n = (C.c_int64 * 8)()
        self.lib.check(self.lib.fn["get_profile"](self.h, ms, n))
        return list(ms), list(n)

    def set_stf_values(self, first_iter: int, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float32)
        self.lib.check(self.lib.fn["set_stf_values"](self.h, C.c_int32(first_iter),
                                                     C.c_int32(v.size), v.ctypes.data_as(_F)))

    def stf_symp(self, first_iter: int, n: int) -> np.ndarray:
        """(n, nstages): the source time function at the sub-stages of steps first_iter+1.. as the
        symplectic loop applies it."""
        out = np.zeros((n, NSTAGES[self.prob.time_scheme]), dtype=np.float32)
        self.lib.check(self.lib.fn["get_stf_symp"](self.h, C.c_int32(first_iter), C.c_int32(n), out.ctypes.data_as(_F)))
        return out

    def set_stream(self, cuda_stream: int):
        self.lib.check(self.lib.fn["set_stream"](self.h, C.c_void_p(cuda_stream)))

    @property
    def iter(self) -> int:
        return int(self.lib.fn["iter"](self.h))

    @property
    def nseismo(self) -> int:
        return int(self.lib.fn["nseismo"](self.h))

    @property
    def nstrain(self) -> int:
        return int(self.lib.fn["nstrain"](self.h))

    @property
    def gpu_launches(self) -> int:
        return int(self.lib.fn["gpu_launches"](self.h))

    def seismograms(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """(nsamples, num_rec, 3) — Fortran recdumpvar slice (3, num_rec, n)."""
        n = self.nseismo - first if n is None else n
        out = np.zeros((n, self.prob.num_rec, 3), dtype=np.float32)
        if n > 0 and self.prob.num_rec > 0:
            self.lib.check(self.lib.fn["fetch_seismograms"](
                self.h, C.c_int32(first), C.c_int32(n), out.ctypes.data_as(_F)))
        return out

    def snapshots(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """(nvars, nsnap, npoints); displ_only: nvars = 3 (s, p, z), see axb_snapshot_layout."""
        n = self.nstrain - first if n is None else n
        npts, nvars = C.c_int32(), C.c_int32()
        self.lib.check(self.lib.fn["snapshot_layout"](self.h, C.byref(npts), C.byref(nvars)))
        out = np.zeros((nvars.value, n, npts.value), dtype=np.float32)
        if n > 0:
            self.lib.check(self.lib.fn["fetch_snapshots"](
                self.h, C.c_int32(first), C.c_int32(n), out.ctypes.data_as(_F)))
        return out

    def xdmf_snapshots(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """(5, nsnap, npoint_plot): u_s, u_p, u_z, straintrace, curlinplane (axb_fetch_xdmf)."""
        cnt = C.c_int32()
        self.lib.check(self.lib.fn["xdmf_count"](self.h, C.byref(cnt)))
        n = cnt.value - first if n is None else n
        out = np.zeros((5, n, self.prob.xdmf["npoint_plot"]), dtype=np.float32)
        if n > 0:
            self.lib.check(self.lib.fn["fetch_xdmf"](self.h, C.c_int32(first), C.c_int32(n), out.ctypes.data_as(_F)))
        return out

    def energy(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        """(n, 4): epot_sol, ekin_sol, epot_flu, ekin_flu sums of this rank for iter first..first+n-1."""
        n = self.iter + 1 - first if n is None else n
        out = np.zeros((n, 4), dtype=np.float32)
        if n > 0:
            self.lib.check(self.lib.fn["fetch_energy"](self.h, C.c_int32(first), C.c_int32(n),
                                                       out.ctypes.data_as(_F)))
        return out

    def _field_shape(self, name: str):
        p = self.prob
        ns, nf = p.mesh.nel_solid, p.mesh.nel_fluid
        if name in ("disp", "velo", "acc0", "acc1"):
            return (3, ns, 5, 5)
        if name in ("chi", "dchi", "ddchi0", "ddchi1"):
            return (nf, 5, 5)
        cg = bool(p.att["coarse_grained"])
        L = int(p.att["n_sls"])
        per = (4,) if cg else (5, 5)
        if name == "memvar":
            return (ns, L, 6) + per
        if name == "src_dev_tm1":
            return (ns, 6) + per
        if name == "src_tr_tm1":
            return (ns,) + per
        raise KeyError(name)

    def get(self, name: str) -> np.ndarray:
        out = np.zeros(self._field_shape(name), dtype=np.float32)
        self.lib.check(self.lib.fn["get_state"](self.h, C.c_int32(FIELDS[name]),
                                                out.ctypes.data_as(_F)))
        return out

    def set(self, name: str, value: np.ndarray):
        v = np.ascontiguousarray(value, dtype=np.float32)
        assert v.shape == self._field_shape(name), (v.shape, self._field_shape(name))
        self.lib.check(self.lib.fn["set_state"](self.h, C.c_int32(FIELDS[name]),
                                                v.ctypes.data_as(_F)))

    def apply_op(self, op: str):
        self.lib.check(self.lib.fn["apply_op"](self.h, C.c_int32(OPS[op])))

    def close(self):
        if self.h:
            self.lib.fn["destroy"](self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def connect_local(lib: Library, loops: Sequence[TimeLoop]):
    arr = (C.c_void_p * len(loops))(*[l.h for l in loops])
    lib.check(lib.fn["connect_local"](arr, C.c_int32(len(loops))))


def run_group(lib: Library, loops: Sequence[TimeLoop], nsteps: int):
    arr = (C.c_void_p * len(loops))(*[l.h for l in loops])
    lib.check(lib.fn["run_group"](arr, C.c_int32(len(loops)), C.c_int32(nsteps)))
