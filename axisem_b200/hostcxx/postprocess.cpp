#include "postprocess.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>

namespace axisem {

void single_simulation_moment(const std::string &t, double amp, double M[6]) {
    for (int k = 0; k < 6; k++) M[k] = 0.0;
    if (t == "mrr") M[0] = amp;
    else if (t == "mtt_p_mpp") { M[1] = amp; M[2] = amp; }
    else if (t == "mtr" || t == "mrt") M[3] = amp;
    else if (t == "mpr" || t == "mrp") M[4] = amp;
    else if (t == "mtp" || t == "mpt") M[5] = amp;
    else if (t == "mtt_m_mpp") { M[1] = amp; M[2] = -amp; }
    else if (t == "explosion") { M[0] = amp; M[1] = amp; M[2] = amp; }
    else throw std::invalid_argument("unknown source type " + t);
}

void moment_from_cmtsolution(const std::string &path, double M[6]) {
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) throw std::runtime_error("cannot open " + path);
    char line[512];
    for (int k = 0; k < 7; k++)
        if (!std::fgets(line, sizeof line, f)) { std::fclose(f); throw std::runtime_error(path + ": short CMTSOLUTION"); }
    for (int k = 0; k < 6; k++) {
        char name[64];
        if (!std::fgets(line, sizeof line, f) || std::sscanf(line, "%63s %lf", name, &M[k]) != 2) {
            std::fclose(f);
            throw std::runtime_error(path + ": cannot read the moment tensor");
        }
        M[k] /= 1.0e7;                                   // dyn cm -> N m
    }
    std::fclose(f);
}

void radiation_prefactor(const std::string &t, const double Mij[6], double magnitude, double lon, double out[3]) {
    double s[6];
    for (int k = 0; k < 6; k++) s[k] = Mij[k] / magnitude;
    if (t == "mrr") { out[0] = s[0]; out[1] = 0.0; out[2] = s[0]; }
    else if (t == "mtt_p_mpp") { out[0] = s[1] + s[2]; out[1] = 0.0; out[2] = s[1] + s[2]; }
    else if (t == "mtr" || t == "mrt" || t == "mpr" || t == "mrp") {
        out[0] = s[3] * std::cos(lon) + s[4] * std::sin(lon);
        out[1] = -s[3] * std::sin(lon) + s[4] * std::cos(lon);
        out[2] = out[0];
    } else if (t == "mtp" || t == "mpt" || t == "mtt_m_mpp") {
        out[0] = (s[1] - s[2]) * std::cos(2.0 * lon) + 2.0 * s[5] * std::sin(2.0 * lon);
        out[1] = (s[2] - s[1]) * std::sin(2.0 * lon) + 2.0 * s[5] * std::cos(2.0 * lon);
        out[2] = out[0];
    } else if (t == "explosion") {
        out[0] = out[1] = out[2] = (s[0] + s[1] + s[2]) / 3.0;
    } else throw std::invalid_argument("unknown source type " + t);
}

void sum_individual_wavefields(std::vector<float> &sum, const float *in, size_t n, const double f[3]) {
    if (sum.size() != 3 * n) sum.assign(3 * n, 0.0f);
    for (size_t k = 0; k < n; k++)
        for (int c = 0; c < 3; c++) sum[3 * k + c] = (float)(sum[3 * k + c] + f[c] * in[3 * k + c]);
}

// rot_mat of Nissen-Meyer, Dahlen, Fournier (GJI 2007), post_processing.F90:749-758
static void rotation_matrix(const SourceLocation &s, double R[3][3]) {
    const double ct = std::cos(s.colat), st = std::sin(s.colat), cp = std::cos(s.lon), sp = std::sin(s.lon);
    R[0][0] = ct * cp; R[0][1] = -sp; R[0][2] = st * cp;
    R[1][0] = ct * sp; R[1][1] = cp;  R[1][2] = st * sp;
    R[2][0] = -st;     R[2][1] = 0.0; R[2][2] = ct;
}

void receiver_location(const SourceLocation &src, double colat, double lon, double &colat_orig, double &lon_orig) {
    const double smallval = 1e-11;
    if (std::fabs(lon - 2.0 * M_PI) < 0.01 * M_PI / 180.0) lon = 0.0;
    const double x0 = std::sin(colat) * std::cos(lon), y0 = std::sin(colat) * std::sin(lon), z0 = std::cos(colat);
    double R[3][3];
    rotation_matrix(src, R);
    double x = R[0][0] * x0 + R[0][1] * y0 + R[0][2] * z0;
    double y = R[1][0] * x0 + R[1][1] * y0 + R[1][2] * z0;
    double z = R[2][0] * x0 + R[2][2] * z0;
    x = std::min(1.0, std::max(-1.0, x)); y = std::min(1.0, std::max(-1.0, y)); z = std::min(1.0, std::max(-1.0, z));
    const double nrm = std::sqrt(x * x + y * y + z * z);
    x /= nrm; y /= nrm; z /= nrm;
    colat_orig = std::acos(z);
    double arg = (x + smallval) / (std::sqrt(x * x + y * y) + smallval);
    arg = std::min(1.0, std::max(-1.0, arg));
    lon_orig = y >= 0.0 ? std::acos(arg) : 2.0 * M_PI - std::acos(arg);
}

void rotate_receiver_comp(const std::string &sys, const SourceLocation &src, double th_rot, double ph_rot,
                          double th_orig, double ph_orig, size_t n, float *seis) {
    const double epsi = 1e-30;                 // epsi_real: "not exactly at the pole"
    double R[3][3];
    rotation_matrix(src, R);
    const bool rotate = src.colat > epsi || src.lon > epsi;
    const double ctr = std::cos(th_rot), str = std::sin(th_rot), cpr = std::cos(ph_rot), spr = std::sin(ph_rot);
    const double cto = std::cos(th_orig), sto = std::sin(th_orig), cpo = std::cos(ph_orig), spo = std::sin(ph_orig);
    for (size_t k = 0; k < n; k++) {
        const double us = seis[3 * k], up = seis[3 * k + 1], uz = seis[3 * k + 2];
        double t[3];
        if (sys == "src") {
            // source-projected frame: to spherical components, no further rotation
            t[0] = ctr * us - str * uz;
            t[1] = up;
            t[2] = str * us + ctr * uz;
        } else {
            // (s, phi, z) -> (x, y, z) of the solver frame, then to the earth-fixed frame
            double v[3] = {cpr * us - spr * up, spr * us + cpr * up, uz};
            if (rotate) {
                for (int a = 0; a < 3; a++) t[a] = R[a][0] * v[0] + R[a][1] * v[1] + R[a][2] * v[2];
            } else {
                t[0] = v[0]; t[1] = v[1]; t[2] = v[2];
            }
        }
        double o[3];
        if (sys == "enz") {
            o[0] = -cto * cpo * t[0] - cto * spo * t[1] + sto * t[2];      // N
            o[1] = -spo * t[0] + cpo * t[1];                               // E
            o[2] = sto * cpo * t[0] + sto * spo * t[1] + cto * t[2];       // Z
        } else if (sys == "sph") {
            o[0] = cto * cpo * t[0] + cto * spo * t[1] - sto * t[2];       // theta
            o[1] = -spo * t[0] + cpo * t[1];                               // phi
            o[2] = sto * cpo * t[0] + sto * spo * t[1] + cto * t[2];       // r
        } else if (sys == "cyl") {
            o[0] = cpo * t[0] + spo * t[1];
            o[1] = -spo * t[0] + cpo * t[1];
            o[2] = t[2];
        } else if (sys == "xyz" || sys == "src") {
            o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
        } else throw std::invalid_argument("unknown receiver component system " + sys);
        for (int c = 0; c < 3; c++) seis[3 * k + c] = (float)o[c];
    }
}

void convolve_with_stf(double t_0, double dt, const std::string &stf, size_t nt, const float *seis, float *out) {
    if (stf != "gauss_0" && stf != "gauss_1") throw std::invalid_argument("convolve_with_stf: gauss_0 or gauss_1");
    size_t N_j = (size_t)(2.0 * POST_SHIFT_FACT1 * t_0 / dt);
    if (N_j > nt) N_j = nt;
    const double alpha = POST_DECAY / t_0, sqrt_pi_inv = 1.0 / std::sqrt(M_PI);
    std::vector<double> src(N_j + 1, 0.0);
    for (size_t j = 1; j <= N_j; j++) {
        const double tau = (double)j * dt;
        if (stf == "gauss_0") {
            const double e = alpha * (tau - POST_SHIFT_FACT1 * t_0);
            src[j] = e < 50.0 ? alpha * std::exp(-e * e) * sqrt_pi_inv / M_PI : 0.0;
        } else {
            double s = -2.0 * alpha * alpha * (tau - POST_SHIFT_FACT1 * t_0) *
                       std::exp(-std::pow(alpha * (tau - POST_SHIFT_FACT1 * t_0), 2));
            src[j] = s / (alpha * std::sqrt(2.0) * std::exp(-2.0));
        }
    }
    for (size_t i = 1; i <= nt; i++)
        for (int c = 0; c < 3; c++) {
            double s = 0.0;
            for (size_t j = 1; j <= N_j && j < i; j++) s += seis[3 * (i - j - 1) + c] * src[j] * dt;
            out[3 * (i - 1) + c] = (float)(s * M_PI);
        }
}

void convolve_gauss(std::vector<float> &x, double dt, double t_0, double decay) {
    const double a = decay / t_0;
    const int half = (int)std::ceil(4.0 * t_0 / dt);
    std::vector<double> g(2 * half + 1);
    for (int k = -half; k <= half; k++) g[k + half] = a / std::sqrt(M_PI) * std::exp(-(a * k * dt) * (a * k * dt)) * dt;
    const int n = (int)x.size();
    std::vector<float> y(n);
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        const int k0 = std::max(-half, i - (n - 1)), k1 = std::min(half, i);
        for (int k = k0; k <= k1; k++) s += g[k + half] * x[i - k];
        y[i] = (float)s;
    }
    x.swap(y);
}

}  // namespace axisem
