#include "time_loop.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>

#include "axisem_b200.h"

namespace axisem {
namespace {

void ck(int rc, const char *what) {
    if (rc != 0) throw SolverError(std::string(what) + ": " + AXB(last_error)());
}
#define CK(call) ck((call), #call)

// the set-up half of the seam: module arrays -> library (the order of include/axisem_b200.h)
axb_handle hand_over(const Modules &m, int device) {
    // mandatory arrays are looked up with at(): a missing one is named in the error, the way
    // the reference fails on an unallocated module array
    auto F = [&](const char *name) { return m.at(name).f32(); };
    auto I = [&](const char *name) { return m.at(name).i32(); };
    axb_handle h = nullptr;
    CK(AXB(create)(&h, device, m.int_of("data_proc%mynum"), m.int_of("data_proc%nproc")));
    const int nel_solid = m.int_of("data_mesh%nel_solid"), nel_fluid = m.int_of("data_mesh%nel_fluid");
    CK(AXB(set_mesh)(h, 4, nel_solid, nel_fluid, m.int_of("data_mesh%nglob_solid"),
                     m.int_of("data_mesh%nglob_fluid"), I("data_mesh%igloc_solid"),
                     I("data_mesh%igloc_fluid"), I("data_mesh%axis_solid"), I("data_mesh%axis_fluid"),
                     I("data_mesh%ax_el_solid"), (int32_t)m.at("data_mesh%ax_el_solid").count(),
                     I("data_mesh%ax_el_fluid"), (int32_t)m.at("data_mesh%ax_el_fluid").count(),
                     F("data_spec%G0"), F("data_spec%G1"), F("data_spec%G1T"),
                     F("data_spec%G2"), F("data_spec%G2T")));
    axb_solid_terms t;
    std::memset(&t, 0, sizeof t);
#define T(n) t.n = m.f("data_matr%" #n)
    T(M11s); T(M21s); T(M41s); T(M12s); T(M22s); T(M32s); T(M42s); T(M11z); T(M21z); T(M41z);
    T(M13s); T(M33s); T(M43s); T(M1phi); T(M2phi); T(M4phi);
    T(M_1); T(M_2); T(M_3); T(M_4); T(M_5); T(M_6); T(M_7); T(M_8);
    T(M_w1); T(M_w2); T(M_w3); T(M_w4); T(M_w5);
    T(M0_w1); T(M0_w2); T(M0_w3); T(M0_w4); T(M0_w5); T(M0_w6); T(M0_w7); T(M0_w8); T(M0_w9); T(M0_w10);
#undef T
    CK(AXB(set_solid_terms)(h, m.int_of("data_source%src_order"), &t));
    if (nel_fluid > 0)
        CK(AXB(set_fluid_terms)(h, F("data_matr%M1chi_fl"), F("data_matr%M2chi_fl"),
                                F("data_matr%M4chi_fl"), m.f("data_matr%M_w_fl"), m.f("data_matr%M0_w_fl"),
                                F("data_matr%inv_mass_fluid"), m.f("data_mesh%fluid_free_surface_mask")));
    CK(AXB(set_mass)(h, F("data_matr%inv_mass_rho")));
    if (m.has("data_matr%unassem_mass_rho_solid"))          // dump_energy
        CK(AXB(set_energy)(h, m.f("data_matr%unassem_mass_rho_solid"), m.f("data_matr%unassem_mass_lam_fluid")));
    if (m.has("data_mesh%solid_absorbing_gamma") || m.has("data_mesh%fluid_absorbing_gamma"))
        CK(AXB(set_sponge)(h, m.f("data_mesh%solid_absorbing_gamma"), m.f("data_mesh%fluid_absorbing_gamma")));
    const int nel_bdry = m.int_of("data_mesh%nel_bdry");
    if (nel_bdry > 0)
        CK(AXB(set_sf_boundary)(h, nel_bdry, I("data_mesh%bdry_solid_el"), I("data_mesh%bdry_fluid_el"),
                                I("data_mesh%bdry_jpol_solid"), I("data_mesh%bdry_jpol_fluid"),
                                F("data_matr%bdry_matr")));
    if (m.int_of("attenuation%anel_true")) {
        axb_attenuation a;
        std::memset(&a, 0, sizeof a);
        a.coarse_grained = m.int_of("attenuation%att_coarse_grained");
        a.n_sls = m.int_of("attenuation%n_sls_attenuation");
        a.do_corr_lowq = m.int_of("attenuation%do_corr_lowq");
        a.y_j = m.at("attenuation%y_j").f64();
        a.exp_w_j_deltat = m.at("attenuation%exp_w_j_deltat").f64();
        a.ts_fac_t = m.at("attenuation%ts_fac_t").f64();
        a.ts_fac_tm1 = m.at("attenuation%ts_fac_tm1").f64();
        a.Q_mu = F("data_matr%Q_mu");
        a.Q_kappa = F("data_matr%Q_kappa");
        a.inv_s_solid = F("data_pointwise%inv_s_solid");
#define A(n) a.n = m.f("data_matr%" #n)
        A(delta_mu_cg4); A(delta_kappa_cg4); A(Y_cg4); A(V_s_eta_cg4); A(V_s_xi_cg4); A(V_z_eta_cg4); A(V_z_xi_cg4);
        A(delta_mu); A(delta_kappa); A(Y); A(V_s_eta); A(V_s_xi); A(V_z_eta); A(V_z_xi);
        A(Y0); A(V0_s_eta); A(V0_s_xi); A(V0_z_eta); A(V0_z_xi);
#undef A
        a.DsDeta_over_J_sol_cg4 = m.f("attenuation%DsDeta_over_J_sol_cg4");
        a.DzDeta_over_J_sol_cg4 = m.f("attenuation%DzDeta_over_J_sol_cg4");
        a.DsDxi_over_J_sol_cg4 = m.f("attenuation%DsDxi_over_J_sol_cg4");
        a.DzDxi_over_J_sol_cg4 = m.f("attenuation%DzDxi_over_J_sol_cg4");
        a.DsDeta_over_J_sol = m.f("data_pointwise%DsDeta_over_J_sol");
        a.DzDeta_over_J_sol = m.f("data_pointwise%DzDeta_over_J_sol");
        a.DsDxi_over_J_sol = m.f("data_pointwise%DsDxi_over_J_sol");
        a.DzDxi_over_J_sol = m.f("data_pointwise%DzDxi_over_J_sol");
        CK(AXB(set_attenuation)(h, &a));
    }
    const Array &stf = m.at("data_source%stf");
    CK(AXB(set_source)(h, m.int_of("data_source%have_src_in_fluid"), m.int_of("data_source%nelsrc"),
                       I("data_source%ielsrc"), F("data_source%source_term_el"), stf.f32(),
                       (int32_t)stf.count()));
    CK(AXB(set_stf_params)(h, m.int_of("data_source%stf_type"), m.real_of("data_source%decay"),
                           m.real_of("data_source%t_0"), m.real_of("data_source%shift_fact"),
                           m.real_of("data_source%magnitude")));
    CK(AXB(set_receivers)(h, m.int_of("data_mesh%num_rec"), I("data_mesh%recfile_el")));
    if (m.int_of("data_io%dump_wavefields", 0))
        CK(AXB(set_kwf)(h, I("data_mesh%kwf_mask"), I("data_mesh%mapping_ijel_ikwf"),
                        m.int_of("data_mesh%npoint_solid_kwf"), m.int_of("data_mesh%npoint_fluid_kwf"),
                        m.f("data_matr%inv_rho_fluid"), m.f("data_pointwise%DsDeta_over_J_flu"),
                        m.f("data_pointwise%DzDeta_over_J_flu"), m.f("data_pointwise%DsDxi_over_J_flu"),
                        m.f("data_pointwise%DzDxi_over_J_flu")));
    if (m.int_of("data_io%dump_wavefields", 0) && m.int_of("data_io%dump_type", 0) != 0)
        CK(AXB(set_dump)(h, m.int_of("data_io%dump_type"), m.int_of("data_io%ibeg"), m.int_of("data_io%iend"),
                         m.int_of("data_io%jbeg"), m.int_of("data_io%jend"),
                         m.f("data_pointwise%DsDeta_over_J_sol"), m.f("data_pointwise%DzDeta_over_J_sol"),
                         m.f("data_pointwise%DsDxi_over_J_sol"), m.f("data_pointwise%DzDxi_over_J_sol"),
                         m.f("data_pointwise%inv_s_solid"), m.f("data_pointwise%inv_s_fluid")));
    if (m.int_of("data_io%dump_xdmf", 0)) {
        const Array &ia = m.at("data_io%i_arr_xdmf"), &ja = m.at("data_io%j_arr_xdmf");
        const bool fl = m.int_of("data_mesh%nel_fluid") > 0;
        CK(AXB(set_xdmf)(h, m.int_of("data_time%snap_it"), (int32_t)ia.count(), (int32_t)ja.count(), ia.i32(), ja.i32(),
                         I("data_mesh%plotting_mask"), I("data_mesh%mapping_ijel_iplot"), m.int_of("data_mesh%npoint_plot"),
                         m.f("data_pointwise%DsDeta_over_J_sol"), m.f("data_pointwise%DzDeta_over_J_sol"),
                         m.f("data_pointwise%DsDxi_over_J_sol"), m.f("data_pointwise%DzDxi_over_J_sol"),
                         m.f("data_pointwise%inv_s_solid"),
                         fl ? m.f("data_pointwise%DsDeta_over_J_flu") : nullptr, fl ? m.f("data_pointwise%DzDeta_over_J_flu") : nullptr,
                         fl ? m.f("data_pointwise%DsDxi_over_J_flu") : nullptr, fl ? m.f("data_pointwise%DzDxi_over_J_flu") : nullptr,
                         fl ? m.f("data_pointwise%inv_s_fluid") : nullptr, fl ? m.f("data_matr%inv_rho_fluid") : nullptr));
    }
    const char *dom[2] = {"solid", "fluid"};
    for (int d = 0; d < 2; d++) {
        const std::string s = dom[d];
        const int nmsg = m.int_of("data_comm%sizerecv_" + s, 0);
        if (nmsg == 0) continue;
        const Array &gl = m.at("data_comm%glocal_index_msg_recv_" + s);     // (maxmsg, nmsg)
        const int maxmsg = gl.dims.size() == 2 ? (int)gl.dims[1] : (int)gl.count() / nmsg;
        CK(AXB(set_halo)(h, d, nmsg, m.i("data_comm%listrecv_" + s), m.i("data_comm%sizemsgrecv_" + s),
                         gl.i32(), maxmsg, m.int_of("data_comm%num_comm_gll_" + s),
                         m.i("data_comm%glob2el_" + s)));
    }
    CK(AXB(set_time)(h, m.int_of("data_time%time_scheme"), m.real_of("data_time%deltat"),
                     m.int_of("data_time%niter"), m.int_of("data_time%seis_it"),
                     m.int_of("data_time%strain_it")));
    CK(AXB(finalize_setup)(h));
    return h;
}

struct Handles {
    std::vector<axb_handle> h;
    ~Handles() { for (axb_handle x : h) if (x) AXB(destroy)(x); }
};

}  // namespace

TimeLoopResult time_loop(const std::vector<Modules> &ranks, const TimeLoopOptions &opt_in, OutputSink *sink) {
    // a zero or negative cadence from the caller would divide by zero in the chunking below
    TimeLoopOptions opt = opt_in;
    opt.check_iter = std::max(1, opt_in.check_iter);
    opt.nc_dumpbuffersize = std::max(1, opt_in.nc_dumpbuffersize);
    if (ranks.empty()) throw SolverError("time_loop: no ranks");
    const int n = (int)ranks.size();
    const Modules &m0 = ranks[0];
    const int niter = m0.int_of("data_time%niter");
    const double deltat = m0.real_of("data_time%deltat");
    const int nsteps = opt.nsteps < 0 ? niter : opt.nsteps;
    if (nsteps > niter) throw SolverError("time_loop: more steps than niter");
    for (const Modules &m : ranks)
        if (m.int_of("data_time%niter") != niter || m.int_of("data_proc%nproc") != m0.int_of("data_proc%nproc"))
            throw SolverError("time_loop: ranks disagree on niter / nproc");

    Handles H;
    for (int r = 0; r < n; r++) H.h.push_back(hand_over(ranks[r], r % std::max(1, opt.ndevices)));
    if (n > 1 || m0.int_of("data_proc%nproc") > 1) CK(AXB(connect_local)(H.h.data(), n));

    if (opt.verbose) {
        std::fprintf(opt.log, "************ S T A R T I N G   T I M E   L O O P *************\n");
        std::fflush(opt.log);
    }
    const int check_it = std::max(1, niter / 20);                 // parameters.F90:932
    const int strain_it = m0.int_of("data_time%strain_it");
    std::vector<int> seis_done(n, 0), snap_done(n, 0);
    std::vector<float> buf;

    auto flush_seis = [&](int r) {
        const int have = AXB(nseismo)(H.h[r]);
        const int num_rec = ranks[r].int_of("data_mesh%num_rec");
        const int cnt = have - seis_done[r];
        if (cnt <= 0 || num_rec == 0) { seis_done[r] = have; return; }
        buf.resize((size_t)3 * num_rec * cnt);
        CK(AXB(fetch_seismograms)(H.h[r], seis_done[r], cnt, buf.data()));
        if (sink) sink->seismograms(ranks[r].int_of("data_proc%mynum"), num_rec, seis_done[r], cnt, buf.data());
        seis_done[r] = have;
    };
    auto flush_snap = [&](int r, bool force) {
        if (!ranks[r].int_of("data_io%dump_wavefields", 0)) return;
        const int have = AXB(nstrain)(H.h[r]);
        const int cnt = have - snap_done[r];
        if (cnt <= 0 || (!force && cnt < opt.nc_dumpbuffersize)) return;
        int32_t np32 = 0, nvars = 3;
        CK(AXB(snapshot_layout)(H.h[r], &np32, &nvars));
        const size_t np = (size_t)np32;
        buf.resize(np * cnt * nvars);
        CK(AXB(fetch_snapshots)(H.h[r], snap_done[r], cnt, buf.data()));
        if (sink) sink->snapshots(ranks[r].int_of("data_proc%mynum"), np, nvars, snap_done[r], cnt, buf.data());
        snap_done[r] = have;
    };

    // chunks end where the host has something to do: progress line, check-point, a full
    // wavefield buffer (a zero or negative cadence from the caller would divide by zero below)

    const auto t0 = std::chrono::steady_clock::now();
    int iter = 0;
    while (iter < nsteps) {
        int next = nsteps;
        next = std::min(next, (iter / opt.check_iter + 1) * opt.check_iter);
        next = std::min(next, (iter / check_it + 1) * check_it);
        if (strain_it > 0) {
            const int span = strain_it * opt.nc_dumpbuffersize;
            next = std::min(next, (iter / span + 1) * span);
        }
        if (n == 1) CK(AXB(run)(H.h[0], next - iter));
        else CK(AXB(run_group)(H.h.data(), n, next - iter));
        iter = next;
        for (int r = 0; r < n; r++) CK(AXB(synchronize)(H.h[r]));   // reports a blow-up like the reference's stop
        if (opt.verbose && iter % opt.check_iter == 0) {
            std::fprintf(opt.log, "  time step:%6d; t=%8.2f s (%5.1f%%)\n", iter, iter * deltat,
                         (double)iter / (double)niter * 100.0);
            std::fflush(opt.log);
        }
        for (int r = 0; r < n; r++) {
            if (iter % check_it == 0 || iter == nsteps) flush_seis(r);
            flush_snap(r, iter == nsteps);
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (sink)
        for (int r = 0; r < n; r++)
            if (ranks[r].has("data_matr%unassem_mass_rho_solid")) {
                buf.resize((size_t)4 * (iter + 1));
                CK(AXB(fetch_energy)(H.h[r], 0, iter + 1, buf.data()));
                sink->energy(ranks[r].int_of("data_proc%mynum"), iter + 1, buf.data());
            }

    if (sink)
        for (int r = 0; r < n; r++)
            if (ranks[r].int_of("data_io%dump_xdmf", 0)) {
                int32_t cnt = 0;
                CK(AXB(xdmf_count)(H.h[r], &cnt));
                const size_t np = (size_t)ranks[r].int_of("data_mesh%npoint_plot");
                buf.resize(std::max<size_t>(np * cnt * 5, 1));
                if (cnt > 0) CK(AXB(fetch_xdmf)(H.h[r], 0, cnt, buf.data()));
                sink->xdmf(ranks[r].int_of("data_proc%mynum"), np, cnt, buf.data());
            }

    TimeLoopResult res;
    res.iter = AXB(iter)(H.h[0]);
    res.nseismo = AXB(nseismo)(H.h[0]);
    res.nstrain = AXB(nstrain)(H.h[0]);
    for (int r = 0; r < n; r++) res.gpu_launches += AXB(gpu_launches)(H.h[r]);
    res.seconds = std::chrono::duration<double>(t1 - t0).count();
    return res;
}

}  // namespace axisem
