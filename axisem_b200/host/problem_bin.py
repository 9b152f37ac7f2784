"""Write one rank's time-loop inputs as an AXBPROB1 container for the native (C++) host.

The C++ host layer (axisem_b200/hostcxx/) is the stand-in for the Fortran side of the seam
`call time_loop` (SOLVER/main.f90:92): it holds the arrays under the names of the Fortran
module variables and hands them to the C ABI.  This writer is the counterpart of
`prepare_waves` leaving those arrays in the modules: every record is named
``<module>%<variable>`` after the reference (data_mesh.f90, data_spec.f90, data_matr.f90,
data_pointwise.f90, data_source.f90, data_time.f90, data_comm.f90, attenuation.f90) and is
stored in exactly the memory order the Fortran holds (column-major, 1-based index values).

Format (little endian): magic "AXBPROB1", u32 nrec, then per record
u16 namelen | name | u8 dtype (0 f32, 1 f64, 2 i32) | u8 ndim | u64 dims[ndim] | u64 nbytes | data.
"""
from __future__ import annotations

import struct

import numpy as np

from ..capi import SCHEMES, SOLID_FIELDS, STF_TYPES, fortran_matrix
from .source import stf_shift

_DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2}


def _rec(out, name, a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    nb = name.encode()
    out.append(struct.pack("<H", len(nb)) + nb + struct.pack("<BB", _DT[a.dtype], a.ndim)
               + b"".join(struct.pack("<Q", d) for d in a.shape) + struct.pack("<Q", a.nbytes))
    out.append(a.tobytes())


def problem_records(p):
    """[(name, array, dtype)] — what crosses the C ABI, under the reference's names."""
    m, b = p.mesh, p.mesh.basis
    f32, f64, i32 = np.float32, np.float64, np.int32
    r = []
    add = lambda n, a, t: r.append((n, a, t))
    add("data_proc%mynum", m.rank, i32)
    add("data_proc%nproc", m.nranks, i32)
    for k in ("nel_solid", "nel_fluid", "nglob_solid", "nglob_fluid", "nel_bdry"):
        add("data_mesh%" + k, int(getattr(m, k)), i32)
    for k in ("igloc_solid", "igloc_fluid", "axis_solid", "axis_fluid", "ax_el_solid", "ax_el_fluid"):
        add("data_mesh%" + k, getattr(m, k), i32)
    add("data_spec%G0", b.G0, f32)
    for k in ("G1", "G1T", "G2", "G2T"):
        add("data_spec%" + k, fortran_matrix(getattr(b, k)), f32)
    add("data_source%src_order", p.src_order, i32)
    for n in SOLID_FIELDS:
        if p.solid.get(n) is not None:
            add("data_matr%" + n, p.solid[n], f32)
    if m.nel_fluid:
        for n in ("M1chi_fl", "M2chi_fl", "M4chi_fl", "M_w_fl", "M0_w_fl"):
            if p.fluid.get(n) is not None:
                add("data_matr%" + n, p.fluid[n], f32)
        add("data_matr%inv_mass_fluid", p.inv_mass_fluid, f32)
        if p.fluid_free_surface_mask is not None:
            add("data_mesh%fluid_free_surface_mask", p.fluid_free_surface_mask, f32)
    add("data_matr%inv_mass_rho", p.inv_mass_rho, f32)
    if getattr(p, "unassem_mass_rho_solid", None) is not None:
        add("data_matr%unassem_mass_rho_solid", p.unassem_mass_rho_solid, f32)
        if m.nel_fluid:
            add("data_matr%unassem_mass_lam_fluid", p.unassem_mass_lam_fluid, f32)
    if p.solid_absorbing_gamma is not None:
        add("data_mesh%solid_absorbing_gamma", p.solid_absorbing_gamma, f32)
    if p.fluid_absorbing_gamma is not None:
        add("data_mesh%fluid_absorbing_gamma", p.fluid_absorbing_gamma, f32)
    if m.nel_bdry:
        for k in ("bdry_solid_el", "bdry_fluid_el", "bdry_jpol_solid", "bdry_jpol_fluid"):
            add("data_mesh%" + k, getattr(m, k), i32)
        add("data_matr%bdry_matr", p.bdry_matr, f32)
    add("attenuation%anel_true", int(bool(p.anel)), i32)
    if p.anel:
        d = p.att
        cg = bool(d["coarse_grained"])
        add("attenuation%att_coarse_grained", int(cg), i32)
        add("attenuation%n_sls_attenuation", int(d["n_sls"]), i32)
        add("attenuation%do_corr_lowq", int(d["do_corr_lowq"]), i32)
        for k in ("y_j", "exp_w_j_deltat", "ts_fac_t", "ts_fac_tm1"):
            add("attenuation%" + k, d[k], f64)
        add("data_matr%Q_mu", d["Q_mu"], f32)
        add("data_matr%Q_kappa", d["Q_kappa"], f32)
        add("data_pointwise%inv_s_solid", p.pw_solid["inv_s"], f32)
        if cg:
            for k in ("delta_mu_cg4", "delta_kappa_cg4"):
                add("data_matr%" + k, d[k], f32)
            for k in ("Y_cg4", "V_s_eta_cg4", "V_s_xi_cg4", "V_z_eta_cg4", "V_z_xi_cg4"):
                add("data_matr%" + k, p.solid[k], f32)
            for k in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
                add(f"attenuation%{k}_over_J_sol_cg4", d[k + "_over_J_cg4"], f32)
        else:
            for k in ("delta_mu", "delta_kappa"):
                add("data_matr%" + k, d[k], f32)
            for k in ("Y", "V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi",
                      "Y0", "V0_s_eta", "V0_s_xi", "V0_z_eta", "V0_z_xi"):
                add("data_matr%" + k, p.solid[k], f32)
            for k in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
                add(f"data_pointwise%{k}_over_J_sol", p.pw_solid[k + "_over_J"], f32)
    add("data_source%have_src_in_fluid", int(bool(getattr(p, "fluid_src", False))), i32)
    add("data_source%nelsrc", int(p.nelsrc), i32)
    add("data_source%ielsrc", p.ielsrc, i32)
    add("data_source%source_term_el", p.source_term_el, f32)
    add("data_source%stf", p.stf, f32)
    s = p.source
    add("data_source%stf_type", STF_TYPES[s.stf_type], i32)
    add("data_source%decay", float(s.decay), f64)
    add("data_source%t_0", float(s.t_0), f64)
    add("data_source%shift_fact", stf_shift(s, p.deltat), f64)
    add("data_source%magnitude", float(s.magnitude), f64)
    add("data_mesh%num_rec", int(p.num_rec), i32)
    add("data_mesh%recfile_el", np.ascontiguousarray(p.recfile_el.T), i32)
    have_kwf = p.kwf is not None and p.strain_it > 0
    add("data_io%dump_wavefields", int(have_kwf), i32)
    if have_kwf:
        q, pf = p.kwf, p.pw_fluid
        add("data_mesh%kwf_mask", q["kwf_mask"], i32)
        add("data_mesh%mapping_ijel_ikwf", q["mapping_ijel_ikwf"], i32)
        add("data_mesh%npoint_solid_kwf", int(q["npoint_solid_kwf"]), i32)
        add("data_mesh%npoint_fluid_kwf", int(q["npoint_fluid_kwf"]), i32)
        add("data_matr%inv_rho_fluid", p.inv_rho_fluid, f32)
        for k in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
            add(f"data_pointwise%{k}_over_J_flu", pf[k + "_over_J"], f32)
        dump_type = getattr(p, "dump_type", "displ_only")
        add("data_io%dump_type", {"displ_only": 0, "strain_only": 1, "fullfields": 2}[dump_type], i32)
        if dump_type != "displ_only":
            ib, ie, jb, je = getattr(p, "dump_block", (0, 4, 0, 4))
            for n_, v_ in (("ibeg", ib), ("iend", ie), ("jbeg", jb), ("jend", je)):
                add("data_io%" + n_, int(v_), i32)
            add("data_pointwise%inv_s_fluid", pf["inv_s"], f32)
            if not (p.anel and not bool(p.att["coarse_grained"])):       # (not written above already)
                for k in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
                    add(f"data_pointwise%{k}_over_J_sol", p.pw_solid[k + "_over_J"], f32)
            if not p.anel:
                add("data_pointwise%inv_s_solid", p.pw_solid["inv_s"], f32)
    x = getattr(p, "xdmf", None)
    add("data_io%dump_xdmf", int(x is not None), i32)
    if x is not None:
        add("data_time%snap_it", int(x["snap_it"]), i32)
        add("data_io%i_arr_xdmf", x["i_arr"], i32)
        add("data_io%j_arr_xdmf", x["j_arr"], i32)
        add("data_mesh%plotting_mask", x["plotting_mask"], i32)
        add("data_mesh%mapping_ijel_iplot", x["mapping_ijel_iplot"], i32)
        add("data_mesh%npoint_plot", int(x["npoint_plot"]), i32)
        have = set(n_ for n_, _, _ in r)
        for k in ("DsDeta", "DzDeta", "DsDxi", "DzDxi"):
            for dom_, pw_ in (("sol", p.pw_solid), ("flu", p.pw_fluid)):
                if f"data_pointwise%{k}_over_J_{dom_}" not in have and k + "_over_J" in pw_:
                    add(f"data_pointwise%{k}_over_J_{dom_}", pw_[k + "_over_J"], f32)
        if "data_pointwise%inv_s_solid" not in have:
            add("data_pointwise%inv_s_solid", p.pw_solid["inv_s"], f32)
        if "data_pointwise%inv_s_fluid" not in have and "inv_s" in p.pw_fluid:
            add("data_pointwise%inv_s_fluid", p.pw_fluid["inv_s"], f32)
        if "data_matr%inv_rho_fluid" not in have:
            add("data_matr%inv_rho_fluid", p.inv_rho_fluid, f32)
    for dom, hs in (("solid", m.halo_solid), ("fluid", m.halo_fluid)):
        add(f"data_comm%sizerecv_{dom}", int(hs.nmsg), i32)
        if hs.nmsg:
            add(f"data_comm%listrecv_{dom}", hs.list_peer, i32)
            add(f"data_comm%sizemsgrecv_{dom}", hs.sizemsg, i32)
            add(f"data_comm%glocal_index_msg_recv_{dom}", hs.glocal_index_msg, i32)
            add(f"data_comm%num_comm_gll_{dom}", int(hs.num_comm_gll), i32)
            add(f"data_comm%glob2el_{dom}", np.ascontiguousarray(hs.glob2el.T), i32)
    add("data_time%time_scheme", SCHEMES[p.time_scheme], i32)
    add("data_time%deltat", float(p.deltat), f64)
    add("data_time%niter", int(p.niter), i32)
    add("data_time%seis_it", int(p.seis_it), i32)
    add("data_time%strain_it", int(p.strain_it), i32)
    return r


# what read_meshdb (hostcxx/meshdb.cpp) provides from the MESHER's database
MESH_LEVEL = ("data_proc%nproc", "data_mesh%nel_solid", "data_mesh%nel_fluid", "data_mesh%nglob_solid",
              "data_mesh%nglob_fluid", "data_mesh%nel_bdry", "data_mesh%igloc_solid", "data_mesh%igloc_fluid",
              "data_mesh%axis_solid", "data_mesh%axis_fluid", "data_mesh%ax_el_solid", "data_mesh%ax_el_fluid",
              "data_spec%G0", "data_spec%G1", "data_spec%G1T", "data_spec%G2", "data_spec%G2T",
              "data_mesh%bdry_solid_el", "data_mesh%bdry_fluid_el", "data_mesh%bdry_jpol_solid",
              "data_mesh%bdry_jpol_fluid")


def save_problem_bin(prob, path: str, without_mesh: bool = False):
    """without_mesh: leave out everything the mesher's database holds (and the halo lists), for
    `axisem_b200_solver terms.axbp+meshdb.datNNNN`."""
    recs = problem_records(prob)
    if without_mesh:
        recs = [r for r in recs if r[0] not in MESH_LEVEL and not r[0].startswith("data_comm%")]
    out = [b"AXBPROB1", struct.pack("<I", len(recs))]
    for name, a, t in recs:
        _rec(out, name, a, t)
    with open(path, "wb") as f:
        f.write(b"".join(out))
