#include "receivers.hpp"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace axisem {
namespace {
const double PI = 3.14159265358979323846;
const double SMALLVAL = 1e-6, SMALLVAL_DBLE = 1e-11;       // global_parameters.f90: smallval (sp build), smallval_dble
}  // namespace

ReceiverList read_receivers_dat(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open " + path);
    long n = -1;
    f >> n;
    if (!f || n < 0) throw std::invalid_argument(path + ": the first line must hold the number of receivers");
    ReceiverList r;
    for (long i = 1; i <= n; i++) {
        double th, ph;
        if (!(f >> th >> ph)) throw std::invalid_argument(path + ": fewer receiver lines than the count on line 1");
        char app[32];
        std::snprintf(app, sizeof app, "%04ld", i);                   // define_io_appendix
        r.name.push_back(std::string("recfile_") + app);
        r.colat_deg.push_back(th);
        r.lon_deg.push_back(ph);
    }
    return r;
}

ReceiverList read_stations(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("cannot open " + path);
    ReceiverList r;
    std::vector<std::string> sta, net;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ls(line);
        std::string name, network;
        double lat, lon, elevation, bury;
        if (!(ls >> name)) continue;                                  // the reference counts readable lines
        if (!(ls >> network >> lat >> lon >> elevation >> bury))
            throw std::invalid_argument(path + ": cannot read 'name network lat lon elevation burial' from: " + line);
        bool seen = false;
        for (size_t k = 0; k < sta.size() && !seen; k++) seen = sta[k] == name && net[k] == network;
        if (seen) { r.redundant.push_back(line); continue; }
        sta.push_back(name);
        net.push_back(network);
        r.name.push_back(name + "_" + network);
        r.colat_deg.push_back(90.0 - lat);
        r.lon_deg.push_back(lon <= 0.0 ? lon + 360.0 : lon);
    }
    return r;
}

void check_receiver_coordinates(const ReceiverList &r) {
    if (r.size() == 0) return;
    double phmin = 1e300, phmax = -1e300, thmax = -1e300;
    for (size_t k = 0; k < r.size(); k++) {
        phmin = std::fmin(phmin, r.lon_deg[k]); phmax = std::fmax(phmax, r.lon_deg[k]);
        thmax = std::fmax(thmax, r.colat_deg[k]);
    }
    if (phmin < 0.0) throw std::invalid_argument("ERROR: We do not allow negative receiver longitudes....");
    if (phmax > 360.001) throw std::invalid_argument("ERROR: We do not allow receiver longitudes larger than 360 degrees....");
    if (thmax < 0.0) throw std::invalid_argument("ERROR: We do not allow negative receiver colatitudes....");
    if (thmax > 180.001) throw std::invalid_argument("ERROR: We do not allow receiver colatitudes larger than 180 degrees....");
}

void rotate_receivers(double srccolat, double srclon, std::vector<double> &colat_deg, std::vector<double> &lon_deg) {
    double rot[3][3];
    rot[0][0] = std::cos(srccolat) * std::cos(srclon); rot[1][1] = std::cos(srclon); rot[2][2] = std::cos(srccolat);
    rot[1][0] = std::cos(srccolat) * std::sin(srclon); rot[2][0] = -std::sin(srccolat); rot[2][1] = 0.0;
    rot[0][1] = -std::sin(srclon); rot[0][2] = std::sin(srccolat) * std::cos(srclon);
    rot[1][2] = std::sin(srccolat) * std::sin(srclon);
    for (auto &row : rot)
        for (double &v : row)
            if (std::fabs(v) < SMALLVAL) v = 0.0;
    for (size_t k = 0; k < colat_deg.size(); k++) {
        const double th = colat_deg[k] * PI / 180.0, ph = lon_deg[k] * PI / 180.0;
        const double x[3] = {std::sin(th) * std::cos(ph), std::sin(th) * std::sin(ph), std::cos(th)};
        double y[3];
        for (int i = 0; i < 3; i++) y[i] = rot[0][i] * x[0] + rot[1][i] * x[1] + rot[2][i] * x[2];     // transpose(rot_mat) x
        const double rr = std::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
        const double c = std::acos(y[2] / (rr + SMALLVAL_DBLE));
        double l = std::acos(y[0] / (rr * std::sin(c) + SMALLVAL_DBLE));
        if (y[1] < 0.0) l = 2 * PI - l;
        colat_deg[k] = c * 180.0 / PI;
        lon_deg[k] = l * 180.0 / PI;
    }
}

void write_receiver_names(const std::string &path, const ReceiverList &r) {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw std::invalid_argument("cannot write " + path);
    for (size_t k = 0; k < r.size(); k++) std::fprintf(f, " %s %.15g %.15g\n", r.name[k].c_str(), r.colat_deg[k], r.lon_deg[k]);
    std::fclose(f);
}

void write_receiver_rotated(const std::string &path, const std::vector<double> &colat_deg, const std::vector<double> &lon_deg) {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw std::invalid_argument("cannot write " + path);
    for (size_t k = 0; k < colat_deg.size(); k++) std::fprintf(f, " %.15g %.15g\n", colat_deg[k], lon_deg[k]);
    std::fclose(f);
}

ReceiverList prepare_receivers(const ReceiverSetup &s, const std::string &prefix, std::vector<double> &colat_deg,
                               std::vector<double> &lon_deg) {
    if (!s.receivers_file.empty() && !s.stations_file.empty())
        throw std::invalid_argument("receivers: give receivers.dat (colatlon) or STATIONS (stations), not both");
    ReceiverList r = s.stations_file.empty() ? read_receivers_dat(s.receivers_file) : read_stations(s.stations_file);
    check_receiver_coordinates(r);
    write_receiver_names(prefix + ".receiver_names.dat", r);
    colat_deg = r.colat_deg;
    lon_deg = r.lon_deg;
    if (s.rot_src()) rotate_receivers((90.0 - s.src_lat_deg) * PI / 180.0, s.src_lon_deg * PI / 180.0, colat_deg, lon_deg);
    write_receiver_rotated(prefix + ".receiver_rotated.dat", colat_deg, lon_deg);
    return r;
}

void write_receiver_pts(const std::string &path, const std::vector<std::vector<int>> &loc2globrec,
                        const std::vector<std::vector<double>> &recfile_th, const std::vector<double> &lon_deg) {
    std::vector<double> th(lon_deg.size(), 0.0);
    std::vector<int> rank(lon_deg.size(), -1);
    for (size_t r = 0; r < loc2globrec.size(); r++)
        for (size_t k = 0; k < loc2globrec[r].size(); k++) {
            const int g = loc2globrec[r][k] - 1;
            if (g < 0 || g >= (int)th.size()) throw std::invalid_argument("receiver_pts: loc2globrec out of range");
            th[g] = recfile_th[r][k];
            rank[g] = (int)r;
        }
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw std::invalid_argument("cannot write " + path);
    for (size_t g = 0; g < th.size(); g++) {
        if (rank[g] < 0) { std::fclose(f); throw std::invalid_argument("PROBLEM: sum of local receivers is different than global!"); }
        std::fprintf(f, " %.15g %.15g %d\n", th[g], lon_deg[g], rank[g]);
    }
    std::fclose(f);
}

}  // namespace axisem
