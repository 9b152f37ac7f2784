// axisem_b200_postproc — solver output -> seismograms in the receiver's component system
// (the job of SOLVER/UTILS/post_processing.F90 for the files axisem_b200_solver writes).
//
//   single simulation (simtype 'single'):
//     axisem_b200_postproc --src mtr [--amplitude 1e20 --magnitude 1e20] --seis RUN.rank0000.seis.f32 ...
//   full moment tensor (simtype 'moment'): the four basis runs and the event's CMTSOLUTION
//     axisem_b200_postproc --cmt CMTSOLUTION --run mrr 1e20 MZZ.seis.f32 --run mtt_p_mpp 1e20 MXX.seis.f32
//                          --run mtr 1e20 MXZ.seis.f32 --run mtp 1e20 MXY.seis.f32 ...
//   run directories in the reference's own layout (USE_NETCDF false; what axisem_b200_solver --rundir and the
//   reference's solver both leave behind): simulation.info, Data/receiver_names.dat, Data/receiver_pts.dat,
//   Data/<receiver>_disp.dat — one directory for simtype 'single', the four of a moment-tensor source with --cmt
//     axisem_b200_postproc --simdir RUN [--simdir ...] [--cmt CMTSOLUTION] --out traces.f32 [--ascii-out DIR]
//   (--ascii-out: DIR/SEISMOGRAMS/<receiver>_disp_post_mij_conv0000_<comp>.dat, "time - shift, value" per line in
//   the reference's (2ES16.7), post_processing.F90:410-425)
//   common:  --stations st.txt --out traces.f32 [--sys enz|sph|cyl|xyz|src] [--srccolat DEG --srclon DEG]
//            [--conv T0 DECAY DT]          zero-phase unit-area Gaussian (comparison with dirac_0 traces)
//            [--stf-conv T0 DT gauss_0|gauss_1]   the reference's causal convolve_with_stf
//
// stations file: one line per receiver of the .seis files, "colat_deg lon_deg" in the solver's
// frame (source at the north pole), as receiver_pts.dat of the reference; .seis is
// recdumpvar(3, num_rec, nseismo); out is (num_rec, 3, nseismo), components in the reference's
// order (enz: N E Z; sph: theta phi r; cyl: s phi z).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "postprocess.hpp"
#include "rundir.hpp"

namespace {
struct Run { std::string type; double magnitude; std::string file; std::vector<float> raw; };

std::vector<float> read_f32(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fseek(f, 0, SEEK_END);
    const long nb = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<float> v((size_t)nb / 4);
    if (std::fread(v.data(), 4, v.size(), f) != v.size()) { std::fclose(f); throw std::runtime_error("short read of " + path); }
    std::fclose(f);
    return v;
}
}  // namespace

int main(int argc, char **argv) {
    std::string sys = "enz", stations, out, cmt, stf_name, ascii_out;
    std::vector<std::string> simdirs;
    std::vector<Run> runs;
    std::string src, seis;
    double amplitude = 1e20, magnitude = 1e20, t_0 = 0, decay = 3.5, dt = 0, stf_t0 = 0, stf_dt = 0;
    axisem::SourceLocation loc;
    for (int k = 1; k < argc; k++) {
        const std::string a = argv[k];
        auto val = [&]() -> const char * { if (k + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", a.c_str()); std::exit(2); } return argv[++k]; };
        if (a == "--src") src = val();
        else if (a == "--seis") seis = val();
        else if (a == "--run") { Run r; r.type = val(); r.magnitude = std::atof(val()); r.file = val(); runs.push_back(r); }
        else if (a == "--cmt") cmt = val();
        else if (a == "--simdir") simdirs.push_back(val());
        else if (a == "--ascii-out") ascii_out = val();
        else if (a == "--sys") sys = val();
        else if (a == "--stations") stations = val();
        else if (a == "--out") out = val();
        else if (a == "--amplitude") amplitude = std::atof(val());
        else if (a == "--magnitude") magnitude = std::atof(val());
        else if (a == "--srccolat") loc.colat = std::atof(val()) * M_PI / 180.0;
        else if (a == "--srclon") loc.lon = std::atof(val()) * M_PI / 180.0;
        else if (a == "--conv") { t_0 = std::atof(val()); decay = std::atof(val()); dt = std::atof(val()); }
        else if (a == "--stf-conv") { stf_t0 = std::atof(val()); stf_dt = std::atof(val()); stf_name = val(); }
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (!src.empty() && !seis.empty()) { Run r; r.type = src; r.magnitude = magnitude; r.file = seis; runs.push_back(r); }
    std::vector<std::string> recnames;
    std::vector<double> sim_colat, sim_lon;
    double shift_fact = 0.0, seis_dt = 0.0;
    try {
        for (const std::string &dir : simdirs) {
            const axisem::SimulationInfo s = axisem::read_simulation_info(dir + "/simulation.info");
            if (s.use_netcdf) throw std::runtime_error(dir + ": the run wrote NetCDF output; this reader takes the ASCII files");
            if (s.src_type2 == "thetaforce" || s.src_type2 == "phiforce")
                throw std::runtime_error("postprocessing for forces with dipole radiation pattern not yet implemented");   // :649-653
            if (recnames.empty()) {
                FILE *f = std::fopen((dir + "/Data/receiver_names.dat").c_str(), "r");
                FILE *g = std::fopen((dir + "/Data/receiver_pts.dat").c_str(), "r");
                if (!f || !g) throw std::runtime_error("cannot open " + dir + "/Data/receiver_names.dat or receiver_pts.dat");
                char name[256], rest[1024];
                double c, l;
                int junk;
                for (int k = 0; k < s.num_rec_tot; k++) {
                    // read(61,*) recname(i): the first item of the line; read(20,*) colat, lon, junk
                    if (std::fscanf(f, "%255s", name) != 1 || !std::fgets(rest, sizeof rest, f) ||
                        std::fscanf(g, "%lf %lf %d", &c, &l, &junk) != 3)
                        throw std::runtime_error(dir + ": fewer receivers than simulation.info says");
                    recnames.push_back(name);
                    sim_colat.push_back(c * M_PI / 180.0);
                    sim_lon.push_back(l * M_PI / 180.0);
                }
                std::fclose(f);
                std::fclose(g);
                loc.colat = s.srccolat; loc.lon = s.srclon;
                shift_fact = s.shift_fact; seis_dt = s.seis_dt;
                amplitude = s.magnitude;
            } else if ((int)recnames.size() != s.num_rec_tot) {
                throw std::runtime_error("PROBLEM with simulation.info parameters in the respective directories");
            }
            Run r;
            r.type = s.src_type2; r.magnitude = s.magnitude; r.file = dir;
            r.raw = axisem::read_disp_files(dir + "/Data", recnames, s.src_type1 == "monopole", s.nseismo);
            runs.push_back(r);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    if (runs.empty() || runs.size() > 4 || (stations.empty() && simdirs.empty()) || out.empty()) {
        std::fprintf(stderr, "usage: axisem_b200_postproc (--src TYPE --seis FILE | --cmt CMTSOLUTION --run TYPE MAGNITUDE FILE ...) "
                             "--stations FILE --out FILE [--sys enz|sph|cyl|xyz|src] [--srccolat DEG --srclon DEG] "
                             "[--amplitude A --magnitude M] [--conv T0 DECAY DT] [--stf-conv T0 DT gauss_0|gauss_1]\n");
        return 2;
    }
    try {
        std::vector<double> colat = sim_colat, lon = sim_lon;
        if (simdirs.empty()) {
            FILE *f = std::fopen(stations.c_str(), "r");
            if (!f) throw std::runtime_error("cannot open " + stations);
            double c, l;
            while (std::fscanf(f, "%lf %lf", &c, &l) == 2) { colat.push_back(c * M_PI / 180.0); lon.push_back(l * M_PI / 180.0); }
            std::fclose(f);
        }
        const size_t nrec = colat.size();
        size_t ns = 0;
        for (Run &r : runs) {
            if (simdirs.empty()) r.raw = read_f32(r.file);
            if (nrec == 0 || r.raw.size() % (3 * nrec) != 0)
                throw std::runtime_error(r.file + " does not hold 3 x num_rec x n values");
            if (ns && r.raw.size() / (3 * nrec) != ns) throw std::runtime_error("runs differ in length");
            ns = r.raw.size() / (3 * nrec);
        }
        double Mij[6];
        if (!cmt.empty()) axisem::moment_from_cmtsolution(cmt, Mij);
        else if (runs.size() == 1) axisem::single_simulation_moment(runs[0].type, amplitude, Mij);
        else throw std::runtime_error("several runs need the event's moment tensor (--cmt)");
        std::vector<float> res(nrec * 3 * ns), one(3 * ns), sum, fil(3 * ns);
        for (size_t r = 0; r < nrec; r++) {
            sum.assign(3 * ns, 0.0f);
            for (const Run &run : runs) {
                double f[3];
                axisem::radiation_prefactor(run.type, Mij, run.magnitude, lon[r], f);
                for (size_t k = 0; k < ns; k++)
                    for (int c = 0; c < 3; c++) one[3 * k + c] = run.raw[c + 3 * (r + nrec * k)];
                axisem::sum_individual_wavefields(sum, one.data(), ns, f);
            }
            if (stf_t0 > 0) {
                axisem::convolve_with_stf(stf_t0, stf_dt, stf_name, ns, sum.data(), fil.data());
                sum = fil;
            }
            double th_orig, ph_orig;
            axisem::receiver_location(loc, colat[r], lon[r], th_orig, ph_orig);
            axisem::rotate_receiver_comp(sys, loc, colat[r], lon[r], th_orig, ph_orig, ns, sum.data());
            for (int c = 0; c < 3; c++) {
                std::vector<float> tr(ns);
                for (size_t k = 0; k < ns; k++) tr[k] = sum[3 * k + c];
                if (t_0 > 0) axisem::convolve_gauss(tr, dt, t_0, decay);
                for (size_t k = 0; k < ns; k++) res[(r * 3 + c) * ns + k] = tr[k];
            }
        }
        FILE *f = std::fopen(out.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write " + out);
        std::fwrite(res.data(), 4, res.size(), f);
        std::fclose(f);
        if (!ascii_out.empty()) {
            if (recnames.empty()) throw std::runtime_error("--ascii-out names the files after the receivers of a --simdir");
            // reccomp is character(len=1) in the reference: 'th' / 'ph' are cut to their first letter (post_processing.F90:40, 116-137)
            static const char *comp[5][3] = {{"N", "E", "Z"}, {"t", "p", "r"}, {"s", "p", "z"}, {"x", "y", "z"}, {"R", "T", "Z"}};
            const int isys = sys == "enz" ? 0 : sys == "sph" ? 1 : sys == "cyl" ? 2 : sys == "xyz" ? 3 : 4;
            axisem::make_directory(ascii_out);
            axisem::make_directory(ascii_out + "/SEISMOGRAMS");
            const int iconv = (int)(stf_t0 > 0 ? stf_t0 : 0.0);
            for (size_t r = 0; r < nrec; r++)
                for (int c = 0; c < 3; c++) {
                    char app[64];
                    std::snprintf(app, sizeof app, "_disp_post_mij_conv%04d_%s.dat", iconv, comp[isys][c]);
                    const std::string path = ascii_out + "/SEISMOGRAMS/" + recnames[r] + app;
                    FILE *g = std::fopen(path.c_str(), "w");
                    if (!g) throw std::runtime_error("cannot write " + path);
                    for (size_t k = 0; k < ns; k++)
                        std::fprintf(g, "%16.7E%16.7E\n", (double)k * seis_dt - shift_fact, (double)res[(r * 3 + c) * ns + k]);
                    std::fclose(g);
                }
        }
        std::printf("%zu receivers x 3 (%s) x %zu samples, %zu run(s)\n", nrec, sys.c_str(), ns, runs.size());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
