// axisem_b200_solver — stand-in for `program axisem` from `call time_loop` on
// (SOLVER/main.f90:92-110): reads one AXBPROB1 container per theta-slice (what read_db +
// prepare_waves leave in the Fortran modules), runs the device-resident time loop for all of
// them in this process (one GPU per slice), and writes the receiver / wavefield buffers the
// reference hands to its NetCDF writers as raw real(4) files:
//   PREFIX.rankNNNN.seis.f32   recdumpvar(3, num_rec, nseismo)      (nc_routines.F90:530-540)
//   PREFIX.rankNNNN.snap.f32   oneddumpvar(npoints, nstrain, 3)     (nc_routines.F90:248,275)
//   PREFIX.rankNNNN.xdmf.f32   xdmf snapshot fields (npoint_plot, nsnap, 5)   (wavefields_io.f90:195-199)
//   PREFIX.info                key = value summary
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "background_models.hpp"
#include "meshdb.hpp"
#include "time_loop.hpp"
#include "precomp.hpp"
#include "receivers.hpp"
#include "rundir.hpp"

namespace {

struct FileSink : axisem::OutputSink {
    struct Rank {
        int num_rec = 0, nseis = 0, nsnap = 0, nvars = 3;
        size_t npoints = 0;
        std::vector<float> seis;                 // (3, num_rec, nseis)
        std::vector<std::vector<float>> snap;    // chunks of (npoints, n, nvars)
        std::vector<int> snap_n;
    };
    std::map<int, Rank> ranks;
    void seismograms(int rank, int num_rec, int first, int n, const float *v) override {
        Rank &r = ranks[rank];
        r.num_rec = num_rec;
        if (first != r.nseis) throw axisem::SolverError("seismogram chunks out of order");
        r.seis.insert(r.seis.end(), v, v + (size_t)3 * num_rec * n);
        r.nseis += n;
    }
    void snapshots(int rank, size_t npoints, int nvars, int first, int n, const float *v) override {
        Rank &r = ranks[rank];
        r.npoints = npoints;
        r.nvars = nvars;
        if (first != r.nsnap) throw axisem::SolverError("snapshot chunks out of order");
        r.snap.emplace_back(v, v + npoints * n * nvars);
        r.snap_n.push_back(n);
        r.nsnap += n;
    }
    std::map<int, std::vector<float>> en;
    void energy(int rank, int n, const float *v) override { en[rank].assign(v, v + (size_t)4 * n); }
    std::map<int, std::vector<float>> xd;
    void xdmf(int rank, size_t npoint_plot, int n, const float *v) override { xd[rank].assign(v, v + npoint_plot * n * 5); }
    void write(const std::string &prefix) const {
        if (!en.empty()) {
            // energy.dat of the reference: t-less table  epot+..., summed over ranks, times two*pi
            const size_t n = en.begin()->second.size() / 4;
            FILE *f = std::fopen((prefix + ".energy.txt").c_str(), "w");
            if (!f) throw axisem::SolverError("cannot write " + prefix + ".energy.txt");
            for (size_t k = 0; k < n; k++) {
                double s[4] = {0, 0, 0, 0};
                for (const auto &kv : en) for (int c = 0; c < 4; c++) s[c] += kv.second[4 * k + c];
                const double tp = 2.0 * 3.14159265358979323846;
                std::fprintf(f, "%zu %.6e %.6e %.6e %.6e %.6e\n", k, tp * s[0], tp * s[1], tp * s[2], tp * s[3],
                             0.5 * tp * (s[0] + s[1] + s[2] + s[3]));
            }
            std::fclose(f);
        }
        for (const auto &kv : xd) {
            // (npoint_plot, nsnap, 5): u_s, u_p, u_z, straintrace, curlinplane — host/xdmf.py:write_xdmf
            // turns it into the reference's xdmf_snap_*.dat + xdmf_xml_NNNN.xdmf
            char app[32];
            std::snprintf(app, sizeof app, ".rank%04d", kv.first);
            FILE *f = std::fopen((prefix + app + ".xdmf.f32").c_str(), "wb");
            if (!f) throw axisem::SolverError("cannot write " + prefix + app + ".xdmf.f32");
            std::fwrite(kv.second.data(), sizeof(float), kv.second.size(), f);
            std::fclose(f);
        }
        for (const auto &kv : ranks) {
            char app[32];
            std::snprintf(app, sizeof app, ".rank%04d", kv.first);
            const Rank &r = kv.second;
            if (r.nseis) {
                FILE *f = std::fopen((prefix + app + ".seis.f32").c_str(), "wb");
                if (!f) throw axisem::SolverError("cannot write " + prefix + app + ".seis.f32");
                std::fwrite(r.seis.data(), sizeof(float), r.seis.size(), f);
                std::fclose(f);
            }
            if (r.nsnap) {
                // reassemble oneddumpvar(npoints, nsnap, nvars) from the buffered chunks
                std::vector<float> all(r.npoints * r.nsnap * r.nvars);
                int off = 0;
                for (size_t c = 0; c < r.snap.size(); c++) {
                    const int n = r.snap_n[c];
                    for (int v = 0; v < r.nvars; v++)
                        std::memcpy(&all[(size_t)v * r.npoints * r.nsnap + (size_t)off * r.npoints],
                                    &r.snap[c][(size_t)v * r.npoints * n], sizeof(float) * r.npoints * n);
                    off += n;
                }
                FILE *f = std::fopen((prefix + app + ".snap.f32").c_str(), "wb");
                if (!f) throw axisem::SolverError("cannot write " + prefix + app + ".snap.f32");
                std::fwrite(all.data(), sizeof(float), all.size(), f);
                std::fclose(f);
            }
        }
    }
};

// what the reference leaves in its run directory with USE_NETCDF false: simulation.info, Data/receiver_names.dat,
// Data/receiver_pts.dat, Data/<receiver>_disp.dat, Data/stf*.dat — the input of its post-processing
void write_rundir(const std::string &dir, const std::vector<axisem::Modules> &ranks, const axisem::PrecompOptions &pre,
                  const axisem::ReceiverSetup &recs, const axisem::ReceiverList &rec_list, const std::vector<double> &rec_lon,
                  const FileSink &sink, const axisem::TimeLoopResult &res) {
    const double PI = 3.14159265358979323846;
    axisem::make_directory(dir);
    axisem::make_directory(dir + "/Data");
    const axisem::Modules &m0 = ranks[0];
    const auto loc2glob = axisem::receiver_indices(ranks);
    const auto th = axisem::receiver_colatitudes(ranks);
    size_t nrec = 0;
    for (const auto &v : loc2glob) nrec += v.size();
    // receivers of the run in the order of the list they came from
    axisem::ReceiverList names = rec_list;
    std::vector<double> lon = rec_lon;
    if (names.size() == 0) {
        for (size_t k = 0; k < nrec; k++) {
            char b[32];
            std::snprintf(b, sizeof b, "recfile_%04zu", k + 1);
            names.name.push_back(b);
            names.colat_deg.push_back(k < pre.rec_colat_deg.size() ? pre.rec_colat_deg[k] : 0.0);
            names.lon_deg.push_back(0.0);
        }
        lon.assign(nrec, 0.0);
    }
    if (names.size() != nrec) throw axisem::SolverError("PROBLEM: sum of local receivers is different than global!");
    axisem::write_receiver_names(dir + "/Data/receiver_names.dat", names);
    axisem::write_receiver_pts(dir + "/Data/receiver_pts.dat", loc2glob, th, lon);
    // seismograms of all ranks in list order
    int nseis = res.nseismo;                       // (rank 0's count; a rank without receivers keeps none)
    for (const auto &kv : sink.ranks) nseis = std::max(nseis, kv.second.nseis);
    std::vector<float> seis((size_t)nseis * nrec * 3, 0.0f);
    for (size_t r = 0; r < ranks.size(); r++) {
        if (loc2glob[r].empty()) continue;
        const auto it = sink.ranks.find(ranks[r].int_of("data_proc%mynum"));
        if (it == sink.ranks.end() || it->second.nseis != nseis) throw axisem::SolverError("rundir: seismograms of a rank are missing");
        const size_t nl = loc2glob[r].size();
        for (int k = 0; k < nseis; k++)
            for (size_t q = 0; q < nl; q++)
                for (int c = 0; c < 3; c++)
                    seis[((size_t)k * nrec + (loc2glob[r][q] - 1)) * 3 + c] = it->second.seis[((size_t)k * nl + q) * 3 + c];
    }
    const int src_order = m0.int_of("data_source%src_order");
    axisem::write_disp_files(dir + "/Data", names.name, src_order == 0, nseis, seis);
    // simulation.info
    axisem::SimulationInfo s;
    {
        const axisem::Array &b = m0.at("data_mesh%bkgrdmodel");
        for (size_t k = 0; k < b.count(); k++) s.bkgrdmodel.push_back((char)b.i32()[k]);
        if (!pre.model.empty()) s.bkgrdmodel = pre.model;
    }
    s.deltat = m0.real_of("data_time%deltat");
    s.niter = m0.int_of("data_time%niter");
    s.src_type1 = src_order == 0 ? "monopole" : (src_order == 1 ? "dipole" : "quadpole");
    s.src_type2 = pre.src_type2;
    s.stf_type = pre.stf_type;
    s.period = pre.t_0;
    s.src_depth_km = pre.src_depth / 1000.0;
    s.srccolat = (90.0 - recs.src_lat_deg) * PI / 180.0;
    s.srclon = recs.src_lon_deg * PI / 180.0;
    s.magnitude = pre.magnitude;
    s.num_rec_tot = (int)nrec;
    s.nseismo = nseis;
    const int seis_it = m0.int_of("data_time%seis_it"), strain_it = m0.int_of("data_time%strain_it");
    s.seis_dt = (double)(float)s.deltat * (double)(float)seis_it;        // real(deltat) * real(seis_it)
    const bool dump = m0.int_of("data_io%dump_wavefields") != 0;
    const double deltat_coarse = s.deltat * (dump ? strain_it : seis_it);
    s.nstrain = dump ? res.nstrain : 0;
    s.strain_dt = dump ? deltat_coarse : 0.0;
    const int snap_it = m0.has("data_time%snap_it") ? m0.int_of("data_time%snap_it") : 0;
    s.nsnap = snap_it > 0 ? s.niter / snap_it : 0;
    s.snap_dt = snap_it > 0 ? s.deltat * snap_it : 0.0;
    s.ibeg = 0; s.iend = 4;                                              // displ_only: the whole element (get_mesh.f90:91-96)
    s.shift_fact = m0.real_of("data_source%shift_fact");
    s.ishift_deltat = (int)std::lround(s.shift_fact / s.deltat);
    s.ishift_seisdt = (int)std::lround(s.shift_fact / (s.deltat * seis_it));
    s.ishift_straindt = (int)std::lround(s.shift_fact / deltat_coarse);
    s.rec_file_type = recs.stations_file.empty() ? "colatlon" : "stations";
    s.nproc = (int)ranks.size();
    // the reference's nelem / nel_fluid are per rank (equal on all ranks): rank 0's
    s.nelem = m0.int_of("data_mesh%nel_solid") + m0.int_of("data_mesh%nel_fluid");
    s.nel_fluid = m0.int_of("data_mesh%nel_fluid");
    axisem::write_simulation_info(dir + "/simulation.info", s);
    // xdmf snapshots (SAVE_SNAPSHOTS with SNAPSHOTS_FORMAT xdmf): iter 0, snap_it, 2 snap_it ...
    for (const axisem::Modules &m : ranks) {
        const int rank = m.int_of("data_proc%mynum");
        const auto it = sink.xd.find(rank);
        if (it == sink.xd.end() || !m.has("data_mesh%npoint_plot")) continue;
        const int npnt = m.int_of("data_mesh%npoint_plot"), nel = m.int_of("data_mesh%nelem_plot");
        if (npnt <= 0) continue;
        const int nsn = (int)(it->second.size() / ((size_t)5 * npnt));
        std::vector<double> times(nsn);
        for (int k = 0; k < nsn; k++) times[k] = (double)k * snap_it * s.deltat;
        axisem::write_xdmf_files(dir + "/Data", rank, npnt, nel, m.at("data_mesh%xdmf_points").f32(), m.i("data_mesh%xdmf_grid"),
                                 it->second.data(), nsn, times, src_order == 0);
    }
    // energy_sol.dat, energy_flu.dat, energy_glob.dat (SAVE_ENERGY; time_evol_wave.F90:103-110, 1514-1523)
    if (!sink.en.empty()) {
        const size_t n = sink.en.begin()->second.size() / 4;
        bool have_fluid = false;
        for (const axisem::Modules &m : ranks) have_fluid = have_fluid || m.int_of("data_mesh%nel_fluid") > 0;
        FILE *fs = have_fluid ? std::fopen((dir + "/Data/energy_sol.dat").c_str(), "w") : nullptr;
        FILE *ff = have_fluid ? std::fopen((dir + "/Data/energy_flu.dat").c_str(), "w") : nullptr;
        FILE *fg = std::fopen((dir + "/Data/energy_glob.dat").c_str(), "w");
        if (!fg || (have_fluid && (!fs || !ff))) throw axisem::SolverError("cannot write " + dir + "/Data/energy_*.dat");
        double t = 0.0;
        for (size_t k = 0; k < n; k++) {
            double e[4] = {0, 0, 0, 0};       // epot_sol, ekin_sol, epot_flu, ekin_flu
            for (const auto &kv : sink.en) for (int c = 0; c < 4; c++) e[c] += kv.second[4 * k + c];
            for (double &v : e) v *= 2.0 * PI;
            if (have_fluid) {
                std::fprintf(fs, "%16.6E%16.6E%16.6E\n", t, e[1], e[0]);
                std::fprintf(ff, "%16.6E%16.6E%16.6E\n", t, e[2], e[3]);
            }
            std::fprintf(fg, "%16.6E%16.6E%16.6E%16.6E\n", t, e[0] + e[2], e[1] + e[3], 0.5 * (e[0] + e[2] + e[1] + e[3]));
            t += s.deltat;
        }
        if (fs) std::fclose(fs);
        if (ff) std::fclose(ff);
        std::fclose(fg);
    }
    // stf.dat, stf_seis.dat, stf_strain.dat (compute_stf, source.f90:186-199)
    if (m0.has("data_source%stf")) {
        const axisem::Array &a = m0.at("data_source%stf");
        const char *fn[3] = {"/Data/stf.dat", "/Data/stf_seis.dat", "/Data/stf_strain.dat"};
        const int every[3] = {1, seis_it, std::max(strain_it, 1)};
        for (int q = 0; q < 3; q++) {
            FILE *f = std::fopen((dir + fn[q]).c_str(), "w");
            if (!f) throw axisem::SolverError("cannot write " + dir + fn[q]);
            for (int i = 1; i <= s.niter && (size_t)i <= a.count(); i++)
                if (i % every[q] == 0) std::fprintf(f, " %16.8E %16.8E\n", (double)((float)i * (float)s.deltat), (double)a.f32()[i - 1]);
            std::fclose(f);
        }
    }
}

void usage() {
    std::fprintf(stderr,
                 "usage: axisem_b200_solver [--steps N] [--devices D] [--dumpbuffer B] [--quiet] [--rundir DIR] --out PREFIX "
                 "rank0.axbp[+meshdb.dat0000] [rank1.axbp[+meshdb.dat0001] ...]\n"
                 "   or: axisem_b200_solver --out PREFIX [--model NAME | --ext-model FILE.bm] [--src TYPE] [--depth KM] [--period T0]\n"
                 "          [--stf gauss_0|gauss_1|gauss_2|errorf|dirac_0|quheavi] [--discrete-choice gaussi|1dirac|...] [--shift SECONDS]\n"
                 "          [--niter N] [--dt DT] [--seis-it K] [--strain-it K] [--attenuation cg4|full] [--scheme NAME]\n"
                 "          [--receivers COLAT,COLAT,... | --receivers-file receivers.dat | --stations STATIONS] [--src-lat DEG --src-lon DEG] [--energy] [--snap-it K]  meshdb.dat0000 [meshdb.dat0001 ...]\n"
                 "       (the second form pre-computes everything from the MESHER's databases, no other input)\n");
}

}  // namespace

int main(int argc, char **argv) {
    axisem::TimeLoopOptions opt;
    std::string prefix;
    std::vector<std::string> files;
    axisem::PrecompOptions pre;
    axisem::ReceiverSetup recs;
    std::string rundir;
    for (int k = 1; k < argc; k++) {
        const std::string a = argv[k];
        auto need = [&](const char *what) -> const char * {
            if (k + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", what); std::exit(2); }
            return argv[++k];
        };
        if (a == "--steps") opt.nsteps = std::atoi(need("--steps"));
        else if (a == "--devices") opt.ndevices = std::atoi(need("--devices"));
        else if (a == "--dumpbuffer") opt.nc_dumpbuffersize = std::max(1, std::atoi(need("--dumpbuffer")));
        else if (a == "--quiet") opt.verbose = false;
        else if (a == "--out") prefix = need("--out");
        else if (a == "--rundir") rundir = need("--rundir");     // the reference's run directory (USE_NETCDF false): rundir.hpp
        else if (a == "--model") pre.model = need("--model");
        else if (a == "--ext-model") {
            try { axisem::set_external_model(axisem::read_external_model(need("--ext-model"))); }
            catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
            pre.model = "external";
        }
        else if (a == "--src") pre.src_type2 = need("--src");
        else if (a == "--depth") pre.src_depth = 1e3 * std::atof(need("--depth"));
        else if (a == "--period") pre.t_0 = std::atof(need("--period"));
        else if (a == "--receivers-file") recs.receivers_file = need("--receivers-file");
        else if (a == "--stations") recs.stations_file = need("--stations");
        else if (a == "--src-lat") recs.src_lat_deg = std::atof(need("--src-lat"));
        else if (a == "--src-lon") recs.src_lon_deg = std::atof(need("--src-lon"));
        else if (a == "--stf") pre.stf_type = need("--stf");
        else if (a == "--discrete-choice") pre.discrete_choice = need("--discrete-choice");
        else if (a == "--shift") pre.shift_seconds = std::atof(need("--shift"));
        else if (a == "--niter") pre.niter = std::atoi(need("--niter"));
        else if (a == "--dt") pre.deltat = std::atof(need("--dt"));
        else if (a == "--seis-it") pre.seis_it = std::atoi(need("--seis-it"));
        else if (a == "--strain-it") { pre.strain_it = std::atoi(need("--strain-it")); pre.dump_wavefields = pre.strain_it > 0; }
        else if (a == "--scheme") pre.time_scheme = need("--scheme");
        else if (a == "--energy") pre.dump_energy = true;
        else if (a == "--snap-it") pre.snap_it = std::atoi(need("--snap-it"));
        else if (a == "--attenuation") { pre.attenuation = true; pre.att.coarse_grained = std::string(need("--attenuation")) != "full"; }
        else if (a == "--receivers") {
            std::string v = need("--receivers");
            size_t pos = 0;
            while (pos < v.size()) {
                size_t c = v.find(',', pos);
                if (c == std::string::npos) c = v.size();
                pre.rec_colat_deg.push_back(std::atof(v.substr(pos, c - pos).c_str()));
                pos = c + 1;
            }
        }
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else files.push_back(a);
    }
    if (files.empty() || prefix.empty()) { usage(); return 2; }
    try {
        std::vector<axisem::Modules> ranks;
        std::vector<double> rec_lon;
        axisem::ReceiverList rec_list;
        const bool from_meshdb = files[0].find(".axbp") == std::string::npos;
        if (from_meshdb) {
            // MESHER databases only: everything else is computed here (precomp.hpp)
            for (size_t r = 0; r < files.size(); r++) ranks.push_back(axisem::read_meshdb(files[r], (int)r));
            if (recs.given()) rec_list = axisem::prepare_receivers(recs, prefix, pre.rec_colat_deg, rec_lon);
            axisem::precompute(ranks, pre);
            if (recs.given()) axisem::write_receiver_pts(prefix + ".receiver_pts.dat", axisem::receiver_indices(ranks), axisem::receiver_colatitudes(ranks), rec_lon);
            const axisem::PrecompChecks c = axisem::precompute_checks(ranks);
            if (opt.verbose)
                std::printf("pre-computation: mass = volume %.10f (solid+fluid over sphere-hollow), S/F boundary term %.10f for %d boundaries\n",
                            (c.solid_volume + c.fluid_volume) / (c.sphere_volume - c.hollow_volume), c.bdry_sum, c.n_sf_boundaries);
        }
        for (const std::string &f : files) {
            if (from_meshdb) break;
            // "terms.axbp+meshdb.datNNNN": mesh-level variables straight from the MESHER's database
            const size_t plus = f.find('+');
            axisem::Modules m = axisem::Modules::read(f.substr(0, plus));
            if (plus != std::string::npos)
                m.merge_missing(axisem::read_meshdb(f.substr(plus + 1), m.int_of("data_proc%mynum")));
            ranks.push_back(std::move(m));
        }
        FileSink sink;
        const axisem::TimeLoopResult res = axisem::time_loop(ranks, opt, &sink);
        sink.write(prefix);
        FILE *f = std::fopen((prefix + ".info").c_str(), "w");
        if (!f) throw axisem::SolverError("cannot write " + prefix + ".info");
        std::fprintf(f, "ranks = %zu\niter = %d\nnseismo = %d\nnstrain = %d\ngpu_launches = %lld\nseconds = %.6f\n",
                     ranks.size(), res.iter, res.nseismo, res.nstrain, (long long)res.gpu_launches, res.seconds);
        for (const auto &kv : sink.ranks)
            std::fprintf(f, "rank%04d = num_rec %d nseis %d npoints %zu nsnap %d\n", kv.first, kv.second.num_rec,
                         kv.second.nseis, kv.second.npoints, kv.second.nsnap);
        std::fclose(f);
        if (!rundir.empty()) {
            if (!from_meshdb) throw axisem::SolverError("--rundir needs the run described on the command line (the meshdb form)");
            write_rundir(rundir, ranks, pre, recs, rec_list, rec_lon, sink, res);
        }
        if (opt.verbose) std::printf("time loop done: %d steps, %.3f s\n", res.iter, res.seconds);
    } catch (const std::exception &e) {
        // the reference writes the message and stops (pcheck / stop)
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
