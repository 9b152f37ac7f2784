#include "rundir.hpp"

#include <sys/stat.h>

#include <cerrno>
#include <cstdarg>
#include <vector>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace axisem {
namespace {

struct Out {
    FILE *f;
    // Fortran edit descriptors: a20 / a45 right-justify (and cut a longer string at the field width)
    static std::string a(const std::string &s, size_t w) { return s.size() >= w ? s.substr(0, w) : std::string(w - s.size(), ' ') + s; }
    void f21(double v, const char *what) { std::fprintf(f, "%22.7f%s\n", v, a(what, 45).c_str()); }
    void f22(long v, const char *what) { std::fprintf(f, "%20ld%s\n", v, a(what, 45).c_str()); }
    void f23(const std::string &v, const char *what) { std::fprintf(f, "%s%s\n", a(v, 20).c_str(), a(what, 45).c_str()); }
    void f24(bool v, const char *what) { std::fprintf(f, "%20s%s\n", v ? "T" : "F", a(what, 45).c_str()); }
    void f25(double v, const char *what) {      // 1pe15.5
        char b[64];
        std::snprintf(b, sizeof b, "%15.5E", v);
        std::fprintf(f, "%s%s\n", b, a(what, 45).c_str());
    }
};

}  // namespace

void make_directory(const std::string &path) {
    if (::mkdir(path.c_str(), 0777) != 0 && errno != EEXIST) throw std::invalid_argument("cannot create directory " + path);
}

void write_simulation_info(const std::string &path, const SimulationInfo &s) {
    Out o{std::fopen(path.c_str(), "w")};
    if (!o.f) throw std::invalid_argument("cannot write " + path);
    o.f23(s.bkgrdmodel, "background model");
    o.f21(s.deltat, "time step [s]");
    o.f22(s.niter, "number of time steps");
    o.f23(s.src_type1, "source type");
    o.f23(s.src_type2, "source type");
    o.f23(s.stf_type, "source time function");
    o.f23(s.simtype, "simtype");
    o.f21(s.period, "dominant source period");
    o.f21(s.src_depth_km, "source depth [km]");
    o.f21(s.srccolat, "Source colatitude");
    o.f21(s.srclon, "Source longitude");
    o.f25(s.magnitude, "scalar source magnitude");
    o.f22(s.num_rec_tot, "number of receivers");
    o.f22(s.nseismo, "length of seismogram [time samples]");
    o.f21(s.seis_dt, "seismogram sampling [s]");
    o.f22(s.nstrain, "number of strain dumps");
    o.f21(s.strain_dt, "strain dump sampling rate [s]");
    o.f22(s.nsnap, "number of snapshot dumps");
    o.f21(s.snap_dt, "snapshot dump sampling rate [s]");
    o.f23(s.rec_comp, "receiver components ");
    o.f22(s.ibeg, "  ibeg: beginning gll index for wavefield dumps");
    o.f22(s.iend, "iend: end gll index for wavefield dumps");
    o.f21(s.shift_fact, "source shift factor [s]");
    o.f22(s.ishift_deltat, "source shift factor for deltat");
    o.f22(s.ishift_seisdt, "source shift factor for seis_dt");
    o.f22(s.ishift_straindt, "source shift factor for deltat_coarse");
    o.f23(s.rec_file_type, "receiver file type");
    o.f21(s.dtheta_rec, "receiver spacing (0 if not even)");
    o.f24(s.use_netcdf, "use netcdf for wavefield output?");
    o.f22(s.nelem, "nelem");
    o.f22(s.nel_fluid, "nel_fluid");
    o.f22(s.nproc, "nproc");
    std::fclose(o.f);
}

SimulationInfo read_simulation_info(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open " + path);
    std::vector<std::string> first;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ls(line);
        std::string tok;
        ls >> tok;
        first.push_back(tok);
    }
    if (first.size() < 32) throw std::invalid_argument(path + ": fewer than the 32 lines of simulation.info");
    auto d = [&](int k) { return std::atof(first[k].c_str()); };
    auto i = [&](int k) { return std::atoi(first[k].c_str()); };
    SimulationInfo s;
    s.bkgrdmodel = first[0]; s.deltat = d(1); s.niter = i(2); s.src_type1 = first[3]; s.src_type2 = first[4];
    s.stf_type = first[5]; s.simtype = first[6]; s.period = d(7); s.src_depth_km = d(8); s.srccolat = d(9); s.srclon = d(10);
    s.magnitude = d(11); s.num_rec_tot = i(12); s.nseismo = i(13); s.seis_dt = d(14); s.nstrain = i(15); s.strain_dt = d(16);
    s.nsnap = i(17); s.snap_dt = d(18); s.rec_comp = first[19]; s.ibeg = i(20); s.iend = i(21); s.shift_fact = d(22);
    s.ishift_deltat = i(23); s.ishift_seisdt = i(24); s.ishift_straindt = i(25); s.rec_file_type = first[26]; s.dtheta_rec = d(27);
    s.use_netcdf = first[28] == "T" || first[28] == ".true." || first[28] == "t";
    s.nelem = i(29); s.nel_fluid = i(30); s.nproc = i(31);
    return s;
}

void write_disp_files(const std::string &data_dir, const std::vector<std::string> &names, bool monopole, int nseis,
                      const std::vector<float> &seis) {
    const size_t nrec = names.size();
    if (seis.size() != (size_t)nseis * nrec * 3) throw std::invalid_argument("write_disp_files: seismogram array of the wrong size");
    for (size_t r = 0; r < nrec; r++) {
        const std::string path = data_dir + "/" + names[r] + "_disp.dat";
        FILE *f = std::fopen(path.c_str(), "w");
        if (!f) throw std::invalid_argument("cannot write " + path);
        for (int k = 0; k < nseis; k++) {
            const float *v = &seis[((size_t)k * nrec + r) * 3];
            if (monopole) std::fprintf(f, " %16.8E %16.8E\n", (double)v[0], (double)v[2]);
            else std::fprintf(f, " %16.8E %16.8E %16.8E\n", (double)v[0], (double)v[1], (double)v[2]);
        }
        std::fclose(f);
    }
}

std::vector<float> read_disp_files(const std::string &data_dir, const std::vector<std::string> &names, bool monopole, int nseis) {
    const size_t nrec = names.size();
    std::vector<float> seis((size_t)nseis * nrec * 3, 0.0f);
    for (size_t r = 0; r < nrec; r++) {
        const std::string path = data_dir + "/" + names[r] + "_disp.dat";
        std::ifstream f(path);
        if (!f) throw std::invalid_argument("cannot open " + path);
        for (int k = 0; k < nseis; k++) {
            float *v = &seis[((size_t)k * nrec + r) * 3];
            double a, b, c = 0.0;
            if (monopole ? !(f >> a >> c) : !(f >> a >> b >> c)) throw std::invalid_argument(path + ": fewer samples than simulation.info says");
            v[0] = (float)a; v[1] = monopole ? 0.0f : (float)b; v[2] = (float)c;
        }
    }
    return seis;
}

namespace {

template <class T>
void put_big_endian(const std::string &path, const T *v, size_t n) {
    static_assert(sizeof(T) == 4, "4-byte items");
    std::vector<unsigned char> b(4 * n);
    for (size_t k = 0; k < n; k++) {
        const unsigned char *p = reinterpret_cast<const unsigned char *>(&v[k]);
        b[4 * k] = p[3]; b[4 * k + 1] = p[2]; b[4 * k + 2] = p[1]; b[4 * k + 3] = p[0];
    }
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::invalid_argument("cannot write " + path);
    std::fwrite(b.data(), 1, b.size(), f);
    std::fclose(f);
}

std::string fmt(const char *f, ...) __attribute__((format(printf, 1, 2)));
std::string fmt(const char *f, ...) {
    char b[512];
    va_list ap;
    va_start(ap, f);
    std::vsnprintf(b, sizeof b, f, ap);
    va_end(ap);
    return b;
}

// the Attribute block of formats 734 / 735 (wavefields_io.f90:286-297)
std::string hyperslab(const char *name, int npoint, int isnap0, int nsnap, const std::string &fname) {
    return fmt("        <Attribute Name=\"%s\" AttributeType=\"Scalar\" Center=\"Node\">\n", name) +
           fmt("            <DataItem ItemType=\"HyperSlab\" Dimensions=\"%10d\" Type=\"HyperSlab\">\n", npoint) +
           "                <DataItem Dimensions=\"3 2\" Format=\"XML\">\n" +
           fmt("                    %10d          0 \n", isnap0) +
           "                             1          1 \n" +
           fmt("                             1 %10d\n", npoint) +
           "                </DataItem>\n" +
           fmt("                <DataItem Dimensions=\"%10d%10d\" NumberType=\"Float\" Format=\"binary\" Endian=\"Big\">\n", nsnap, npoint) +
           "                   " + fname + "\n" +
           "                </DataItem>\n            </DataItem>\n        </Attribute>\n";
}

}  // namespace

void write_xdmf_files(const std::string &dir, int rank, int npnt, int nel, const float *points, const int32_t *grid,
                      const float *fields, int nsnap, const std::vector<double> &times, bool monopole) {
    const std::string app = fmt("%04d", rank);
    put_big_endian(dir + "/xdmf_points_" + app + ".dat", points, (size_t)2 * npnt);
    put_big_endian(dir + "/xdmf_grid_" + app + ".dat", grid, (size_t)4 * nel);
    const char *names[5] = {"s", "p", "z", "trace", "curlip"};
    for (int v = 0; v < 5; v++) {
        if (v == 1 && monopole) continue;                          // unit 13101 is not opened
        put_big_endian(dir + "/xdmf_snap_" + names[v] + "_" + app + ".dat", fields + (size_t)v * nsnap * npnt, (size_t)nsnap * npnt);
    }
    const std::string head = "<?xml version=\"1.0\" ?>\n<!DOCTYPE Xdmf SYSTEM \"Xdmf.dtd\" []>\n"
                             "<Xdmf xmlns:xi=\"http://www.w3.org/2003/XInclude\" Version=\"2.2\">\n<Domain>\n";
    {
        std::string s = head + "<Grid Name=\"CellsTime\" GridType=\"Collection\" CollectionType=\"Temporal\">\n"
                               "  <Grid GridType=\"Uniform\">\n    <Time Value=\"0.000\" />\n" +
                        fmt("    <Topology TopologyType=\"Quadrilateral\" NumberOfElements=\"%10d\">\n", nel) +
                        fmt("      <DataItem Dimensions=\"%10d 4\" NumberType=\"Int\" Format=\"binary\" Endian=\"Big\">\n", nel) +
                        "        xdmf_grid_" + app + ".dat\n      </DataItem>\n    </Topology>\n"
                        "    <Geometry GeometryType=\"XY\">\n" +
                        fmt("      <DataItem Dimensions=\"%10d 2\" NumberType=\"Float\" Format=\"binary\" Endian=\"Big\">\n", npnt) +
                        "        xdmf_points_" + app + ".dat\n      </DataItem>\n    </Geometry>\n"
                        "  </Grid>\n</Grid>\n</Domain>\n</Xdmf>\n";
        std::ofstream f(dir + "/xdmf_meshonly_" + app + ".xdmf");
        if (!f) throw std::invalid_argument("cannot write " + dir + "/xdmf_meshonly_" + app + ".xdmf");
        f << s;
    }
    std::string xml = head + "\n" +
                      fmt("<DataItem Name=\"grid\" Dimensions=\"%10d 4\" NumberType=\"Int\" Format=\"binary\" Endian=\"Big\">\n", nel) +
                      "  xdmf_grid_" + app + ".dat\n</DataItem>\n" +
                      fmt("<DataItem Name=\"points\" Dimensions=\"%10d 2\" NumberType=\"Float\" Format=\"binary\" Endian=\"Big\">\n", npnt) +
                      "  xdmf_points_" + app + ".dat\n</DataItem>\n\n"
                      "<Grid Name=\"CellsTime\" GridType=\"Collection\" CollectionType=\"Temporal\">\n\n";
    const int ncomp = monopole ? 2 : 3;
    const char *comps[3] = {"u_s", monopole ? "u_z" : "u_p", "u_z"};
    const char *files[3] = {"s", monopole ? "z" : "p", "z"};
    for (int k = 0; k < nsnap; k++) {
        const std::string sn = fmt("%04d", k + 1);
        xml += "    <Grid Name=\"" + sn + "\" GridType=\"Uniform\">\n" +
               fmt("        <Time Value=\"%8.2f\" />\n", times[k]) +
               fmt("        <Topology TopologyType=\"Quadrilateral\" NumberOfElements=\"%10d\">\n", nel) +
               "            <DataItem Reference=\"XML\">\n                /Xdmf/Domain/DataItem[@Name=\"grid\"]\n"
               "            </DataItem>\n        </Topology>\n        <Geometry GeometryType=\"XY\">\n"
               "            <DataItem Reference=\"XML\">\n                /Xdmf/Domain/DataItem[@Name=\"points\"]\n"
               "            </DataItem>\n        </Geometry>\n";
        for (int c = 0; c < ncomp; c++) xml += hyperslab(comps[c], npnt, k, nsnap, std::string("xdmf_snap_") + files[c] + "_" + app + ".dat");
        std::string terms, refs;
        for (int c = 0; c < ncomp; c++) {
            terms += fmt("%s$%d * $%d", c ? " + " : "", c, c);
            refs += "                <DataItem Reference=\"XML\">\n"
                    "                    /Xdmf/Domain/Grid[@Name=\"CellsTime\"]/Grid[@Name=\"" + sn + "\"]/Attribute[@Name=\"" + comps[c] +
                    "\"]/DataItem[1]\n                </DataItem>\n";
        }
        xml += "        <Attribute Name=\"abs\" AttributeType=\"Scalar\" Center=\"Node\">\n"
               "            <DataItem ItemType=\"Function\" Function=\"sqrt(" + terms + ")\" Dimensions=\"" + fmt("%10d", npnt) + "\">\n" + refs +
               "            </DataItem>\n        </Attribute>\n";
        xml += hyperslab("straintrace", npnt, k, nsnap, "xdmf_snap_trace_" + app + ".dat");
        xml += hyperslab("curlinplane", npnt, k, nsnap, "xdmf_snap_curlip_" + app + ".dat");
        xml += "    </Grid>\n\n";
    }
    xml += "</Grid>\n</Domain>\n</Xdmf>\n";
    std::ofstream f(dir + "/xdmf_xml_" + app + ".xdmf");
    if (!f) throw std::invalid_argument("cannot write " + dir + "/xdmf_xml_" + app + ".xdmf");
    f << xml;
}

}  // namespace axisem
