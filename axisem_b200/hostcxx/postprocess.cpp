#include "postprocess.hpp"

#include <cmath>
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

void rotate_receiver_comp(const std::string &sys, double colat, size_t n, const float *spz, float *out) {
    const double st = std::sin(colat), ct = std::cos(colat);
    for (size_t k = 0; k < n; k++) {
        const double us = spz[3 * k], up = spz[3 * k + 1], uz = spz[3 * k + 2];
        const double ur = us * st + uz * ct;          // radial, up
        const double ut = us * ct - uz * st;          // colatitudinal, south
        if (sys == "enz") { out[3 * k] = (float)up; out[3 * k + 1] = (float)(-ut); out[3 * k + 2] = (float)ur; }
        else if (sys == "sph") { out[3 * k] = (float)ur; out[3 * k + 1] = (float)ut; out[3 * k + 2] = (float)up; }
        else if (sys == "cyl") { out[3 * k] = (float)us; out[3 * k + 1] = (float)up; out[3 * k + 2] = (float)uz; }
        else throw std::invalid_argument("unknown receiver component system " + sys);
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
