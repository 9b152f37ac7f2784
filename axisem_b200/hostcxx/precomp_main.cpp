// axisem_b200_precomp — MESHER databases -> complete time-loop inputs (module variables), natively.
//   axisem_b200_precomp --out PREFIX [--model NAME | --ext-model FILE.bm] [--src TYPE] [--depth KM] [--period T0]
//        [--niter N] [--dt DT] [--seis-it K] [--strain-it K] [--attenuation cg4|full] [--scheme NAME]
//        [--receivers COLAT,...] [--energy] meshdb.dat0000 [meshdb.dat0001 ...]
// writes PREFIX.rankNNNN.axbp (what axisem_b200_solver takes) and prints the reference's
// self-checks of the pre-computation (mass = volume, S/F boundary term = 2 per boundary).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cmath>

#include "background_models.hpp"
#include "mapping.hpp"
#include "meshdb.hpp"
#include "precomp.hpp"
#include "receivers.hpp"

// --mapping-check: the four element mappings on one spherical-shell element (20-31 degrees,
// 3000-3600 km): corners, derivatives against central differences, the defining geometry of the
// semi-analytic elements
static int mapping_check() {
    const double th_a = 20.0 * M_PI / 180, th_b = 31.0 * M_PI / 180, r_a = 3.0e6, r_b = 3.6e6;
    const double th_m = 0.5 * (th_a + th_b), r_m = 0.5 * (r_a + r_b);
    const double th[8] = {th_a, th_m, th_b, th_b, th_b, th_m, th_a, th_a};
    const double r[8] = {r_a, r_a, r_a, r_m, r_b, r_b, r_b, r_m};
    double nodes[8][2];
    for (int k = 0; k < 8; k++) { nodes[k][0] = r[k] * std::sin(th[k]); nodes[k][1] = r[k] * std::cos(th[k]); }
    const char *names[4] = {"curved", "linear", "semino", "semiso"};
    const double cx[4] = {-1, 1, 1, -1}, ce[4] = {-1, -1, 1, 1};
    const int cn[4] = {0, 2, 4, 6};
    for (int t = 0; t < 4; t++) {
        double corner = 0.0, derr = 0.0, dmax = 0.0;
        for (int c = 0; c < 4; c++) {
            const axisem::MapPoint p = axisem::map_element(t, nodes, cx[c], ce[c], 0.0);
            corner = std::fmax(corner, std::hypot(p.s - nodes[cn[c]][0], p.z - nodes[cn[c]][1]));
        }
        for (double xi : {-0.9, -0.3, 0.2, 0.8})
            for (double eta : {-0.8, -0.1, 0.5, 0.95}) {
                const double h = 1e-6;
                const axisem::MapPoint p = axisem::map_element(t, nodes, xi, eta, 0.0);
                const axisem::MapPoint xp = axisem::map_element(t, nodes, xi + h, eta, 0.0), xm = axisem::map_element(t, nodes, xi - h, eta, 0.0);
                const axisem::MapPoint ep = axisem::map_element(t, nodes, xi, eta + h, 0.0), em = axisem::map_element(t, nodes, xi, eta - h, 0.0);
                const double fd[4] = {(xp.s - xm.s) / (2 * h), (xp.z - xm.z) / (2 * h), (ep.s - em.s) / (2 * h), (ep.z - em.z) / (2 * h)};
                const double an[4] = {p.dsdxi, p.dzdxi, p.dsdeta, p.dzdeta};
                for (int k = 0; k < 4; k++) { derr = std::fmax(derr, std::fabs(fd[k] - an[k])); dmax = std::fmax(dmax, std::fabs(an[k])); }
            }
        std::printf("%s_corner_err %.3e\n%s_derivative_err %.3e\n", names[t], corner, names[t], derr / dmax);
        if (t >= 2) {
            const bool top_curved = t == 2;
            const int la = top_curved ? 0 : 6, lb = top_curved ? 2 : 4;
            const double Rc = top_curved ? r_b : r_a;
            double line = 0.0, ell = 0.0;
            for (double xi : {-1.0, -0.5, 0.0, 0.4, 1.0}) {
                const axisem::MapPoint pl = axisem::map_element(t, nodes, xi, top_curved ? -1.0 : 1.0, 0.0);
                const axisem::MapPoint pc = axisem::map_element(t, nodes, xi, top_curved ? 1.0 : -1.0, 0.0);
                const double dx = nodes[lb][0] - nodes[la][0], dz = nodes[lb][1] - nodes[la][1];
                line = std::fmax(line, std::fabs((pl.s - nodes[la][0]) * dz - (pl.z - nodes[la][1]) * dx) / std::hypot(dx, dz));
                ell = std::fmax(ell, std::fabs(std::hypot(pc.s, pc.z) / Rc - 1.0));
            }
            std::printf("%s_line_err %.3e\n%s_ellipse_err %.3e\n", names[t], line, names[t], ell);
        }
    }
    return 0;
}

// prints "r_km idom rho vpv vsv vph vsh eta qmu qka" per radius, and the model's discontinuities first
int model_eval(const std::string &name, const std::string &list) {
    try {
        if (name == "external") {
            const axisem::ExternalModel &E = axisem::external_model();
            std::printf("ndisc %d anelastic %d anisotropic %d name %s\n", E.ndisc(), (int)E.anelastic, (int)E.anisotropic, E.name.c_str());
            for (int k = 1; k <= E.ndisc(); k++) std::printf("discont %d %.6f fluid %d\n", k, E.discont(k) / 1000.0, (int)E.fluid(k));
        } else {
            const auto &d = axisem::model_domains(name);
            std::printf("ndisc %d anelastic %d anisotropic %d name %s\n", (int)d.size(), (int)axisem::model_is_anelastic(name),
                        (int)axisem::model_is_ani(name), name.c_str());
            for (size_t k = 0; k < d.size(); k++) std::printf("discont %d %.6f fluid %d\n", (int)k + 1, d[k].r_top_km, (int)d[k].fluid);
        }
        size_t pos = 0;
        while (pos < list.size()) {
            size_t c = list.find(',', pos);
            if (c == std::string::npos) c = list.size();
            std::string tok = list.substr(pos, c - pos);
            pos = c + 1;
            bool upper = false;
            if (!tok.empty() && tok.back() == '+') { upper = true; tok.pop_back(); }
            const double r = 1000.0 * std::atof(tok.c_str());
            int idom;
            if (name == "external") {
                const axisem::ExternalModel &E = axisem::external_model();
                idom = E.ndisc();
                for (int k = 1; k <= E.ndisc(); k++) {      // from the surface: first domain that holds r
                    const double bot = k < E.ndisc() ? E.discont(k + 1) : 0.0;
                    if (r <= E.discont(k) && (upper ? r >= bot : r > bot)) { idom = k; break; }
                }
            } else idom = axisem::model_domain_of(name, r, upper);
            const axisem::ModelValues v = axisem::model_evaluate(name, r, idom);
            std::printf("%.6f %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", r / 1000.0, idom, v.rho, v.vpv, v.vsv, v.vph, v.vsh, v.eta,
                        v.qmu, v.qkappa);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}

int main(int argc, char **argv) {
    axisem::PrecompOptions pre;
    axisem::ReceiverSetup recs;
    std::string prefix;
    std::vector<std::string> files;
    for (int k = 1; k < argc; k++) {
        const std::string a = argv[k];
        auto need = [&]() -> const char * { if (k + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", a.c_str()); std::exit(2); } return argv[++k]; };
        if (a == "--mapping-check") return mapping_check();
        else if (a == "--out") prefix = need();
        else if (a == "--model") pre.model = need();
        else if (a == "--ext-model") {        // bkgrdmodel = 'external': the tabulated .bm file of the run
            try { axisem::set_external_model(axisem::read_external_model(need())); }
            catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
            pre.model = "external";
        }
        else if (a == "--model-eval") {       // NAME R_KM[+][,R_KM[+]...]: the model at these radii ('+': upper side of a discontinuity)
            const std::string name = need(), v = need();
            return model_eval(name, v);
        }
        else if (a == "--receivers-file") recs.receivers_file = need();      // receivers.dat (RECFILE_TYPE colatlon)
        else if (a == "--stations") recs.stations_file = need();             // STATIONS (RECFILE_TYPE stations)
        else if (a == "--src-lat") recs.src_lat_deg = std::atof(need());     // SOURCE_LAT / SOURCE_LON [deg]
        else if (a == "--src-lon") recs.src_lon_deg = std::atof(need());
        else if (a == "--src") pre.src_type2 = need();
        else if (a == "--depth") pre.src_depth = 1e3 * std::atof(need());
        else if (a == "--period") pre.t_0 = std::atof(need());
        else if (a == "--stf") pre.stf_type = need();
        else if (a == "--discrete-choice") pre.discrete_choice = need();
        else if (a == "--shift") pre.shift_seconds = std::atof(need());
        else if (a == "--magnitude") pre.magnitude = std::atof(need());
        else if (a == "--niter") pre.niter = std::atoi(need());
        else if (a == "--dt") pre.deltat = std::atof(need());
        else if (a == "--seis-it") pre.seis_it = std::atoi(need());
        else if (a == "--strain-it") { pre.strain_it = std::atoi(need()); pre.dump_wavefields = pre.strain_it > 0; }
        else if (a == "--scheme") pre.time_scheme = need();
        else if (a == "--energy") pre.dump_energy = true;
        else if (a == "--snap-it") pre.snap_it = std::atoi(need());
        else if (a == "--xdmf-region") {     // RMIN_KM RMAX_KM COLAT_MIN_DEG COLAT_MAX_DEG
            pre.xdmf_rmin = 1e3 * std::atof(need()); pre.xdmf_rmax = 1e3 * std::atof(need());
            pre.xdmf_thetamin = std::atof(need()) * 3.14159265358979323846 / 180.0;
            pre.xdmf_thetamax = std::atof(need()) * 3.14159265358979323846 / 180.0;
        }
        else if (a == "--attenuation") { pre.attenuation = true; pre.att.coarse_grained = std::string(need()) != "full"; }
        else if (a == "--fit-sls") {          // N F_MIN F_MAX SEED: NR_LIN_SOLIDS, F_MIN, F_MAX of inparam_advanced
            const int n = std::atoi(need());
            const double f0 = std::atof(need()), f1 = std::atof(need());
            const uint64_t seed = (uint64_t)std::atoll(need());
            if (n < 1 || n > 8 || !(f0 > 0 && f1 > f0)) { std::fprintf(stderr, "--fit-sls N F_MIN F_MAX SEED: 1 <= N <= 8, 0 < F_MIN < F_MAX\n"); return 2; }
            const double chi = axisem::fit_linear_solids(pre.att, n, f0, f1, seed);
            std::printf("sls_misfit %.6e\n", chi);
        }
        else if (a == "--receivers") {
            const std::string v = need();
            size_t pos = 0;
            while (pos < v.size()) {
                size_t c = v.find(',', pos);
                if (c == std::string::npos) c = v.size();
                pre.rec_colat_deg.push_back(std::atof(v.substr(pos, c - pos).c_str()));
                pos = c + 1;
            }
        } else files.push_back(a);
    }
    if (files.empty() || prefix.empty()) {
        std::fprintf(stderr, "usage: axisem_b200_precomp --out PREFIX [options] meshdb.dat0000 [meshdb.dat0001 ...]\n");
        return 2;
    }
    try {
        std::vector<axisem::Modules> ranks;
        for (size_t r = 0; r < files.size(); r++) ranks.push_back(axisem::read_meshdb(files[r], (int)r));
        std::vector<double> rec_lon;
        if (recs.given()) axisem::prepare_receivers(recs, prefix, pre.rec_colat_deg, rec_lon);
        axisem::precompute(ranks, pre);
        if (recs.given()) axisem::write_receiver_pts(prefix + ".receiver_pts.dat", axisem::receiver_indices(ranks), axisem::receiver_colatitudes(ranks), rec_lon);
        const axisem::PrecompChecks c = axisem::precompute_checks(ranks);
        std::printf("mass_over_volume %.12f\nbdry_sum %.12f\nn_sf_boundaries %d\n",
                    (c.solid_volume + c.fluid_volume) / (c.sphere_volume - c.hollow_volume), c.bdry_sum, c.n_sf_boundaries);
        for (size_t r = 0; r < ranks.size(); r++) {
            char app[32];
            std::snprintf(app, sizeof app, ".rank%04zu.axbp", r);
            axisem::write_container(ranks[r], prefix + app);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
