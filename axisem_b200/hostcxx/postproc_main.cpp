// axisem_b200_postproc — solver output -> seismograms in the receiver's component system.
//   axisem_b200_postproc --src mtr [--amplitude 1e20 --magnitude 1e20] --sys enz [--conv T0 DECAY DT]
//                        --stations st.txt --seis RUN.rank0000.seis.f32 --out traces.f32
// stations file: one line per receiver of the .seis file, "colat_deg lon_deg"; .seis is
// recdumpvar(3, num_rec, nseismo) as axisem_b200_solver writes it; out is (num_rec, 3, nseismo).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "postprocess.hpp"

int main(int argc, char **argv) {
    std::string src, sys = "enz", stations, seis, out;
    double amplitude = 1e20, magnitude = 1e20, t_0 = 0, decay = 3.5, dt = 0;
    for (int k = 1; k < argc; k++) {
        const std::string a = argv[k];
        auto val = [&]() -> const char * { if (k + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", a.c_str()); std::exit(2); } return argv[++k]; };
        if (a == "--src") src = val();
        else if (a == "--sys") sys = val();
        else if (a == "--stations") stations = val();
        else if (a == "--seis") seis = val();
        else if (a == "--out") out = val();
        else if (a == "--amplitude") amplitude = std::atof(val());
        else if (a == "--magnitude") magnitude = std::atof(val());
        else if (a == "--conv") { t_0 = std::atof(val()); decay = std::atof(val()); dt = std::atof(val()); }
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (src.empty() || stations.empty() || seis.empty() || out.empty()) {
        std::fprintf(stderr, "usage: axisem_b200_postproc --src TYPE --stations FILE --seis FILE --out FILE [--sys enz|sph|cyl] "
                             "[--amplitude A --magnitude M] [--conv T0 DECAY DT]\n");
        return 2;
    }
    try {
        std::vector<double> colat, lon;
        {
            FILE *f = std::fopen(stations.c_str(), "r");
            if (!f) throw std::runtime_error("cannot open " + stations);
            double c, l;
            while (std::fscanf(f, "%lf %lf", &c, &l) == 2) { colat.push_back(c * M_PI / 180.0); lon.push_back(l * M_PI / 180.0); }
            std::fclose(f);
        }
        const size_t nrec = colat.size();
        std::vector<float> raw;
        {
            FILE *f = std::fopen(seis.c_str(), "rb");
            if (!f) throw std::runtime_error("cannot open " + seis);
            std::fseek(f, 0, SEEK_END);
            const long nb = std::ftell(f);
            std::fseek(f, 0, SEEK_SET);
            raw.resize((size_t)nb / 4);
            if (std::fread(raw.data(), 4, raw.size(), f) != raw.size()) throw std::runtime_error("short read");
            std::fclose(f);
        }
        if (nrec == 0 || raw.size() % (3 * nrec) != 0) throw std::runtime_error("seismogram file does not hold 3 x num_rec x n values");
        const size_t ns = raw.size() / (3 * nrec);
        double Mij[6];
        axisem::single_simulation_moment(src, amplitude, Mij);
        std::vector<float> res(nrec * 3 * ns), spz(3 * ns), rot(3 * ns);
        for (size_t r = 0; r < nrec; r++) {
            double f[3];
            axisem::radiation_prefactor(src, Mij, magnitude, lon[r], f);
            for (size_t k = 0; k < ns; k++)
                for (int c = 0; c < 3; c++) spz[3 * k + c] = (float)(f[c] * raw[c + 3 * (r + nrec * k)]);
            axisem::rotate_receiver_comp(sys, colat[r], ns, spz.data(), rot.data());
            for (int c = 0; c < 3; c++) {
                std::vector<float> tr(ns);
                for (size_t k = 0; k < ns; k++) tr[k] = rot[3 * k + c];
                if (t_0 > 0) axisem::convolve_gauss(tr, dt, t_0, decay);
                for (size_t k = 0; k < ns; k++) res[(r * 3 + c) * ns + k] = tr[k];
            }
        }
        FILE *f = std::fopen(out.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write " + out);
        std::fwrite(res.data(), 4, res.size(), f);
        std::fclose(f);
        std::printf("%zu receivers x 3 (%s) x %zu samples\n", nrec, sys.c_str(), ns);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
