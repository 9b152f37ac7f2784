#include "mapping.hpp"

#include <cmath>
#include <stdexcept>

namespace axisem {
namespace {

const double PI = 3.14159265358979323846;

// ---- curved: concentric / spheroidal elements (analytic_spheroid_mapping.f90) -----------------
void theta_r(const double n[8][2], double th[8], double r[8]) {
    const double min_distance_nondim = 1e-10;        // get_mesh.f90: of the order of the mesher's value
    for (int i = 0; i < 8; i++) {
        r[i] = std::sqrt(n[i][0] * n[i][0] + n[i][1] * n[i][1]);
        th[i] = r[i] != 0.0 ? std::acos(n[i][1] / r[i]) : 0.0;
        if (th[i] < PI * min_distance_nondim) th[i] = 0.0;
        if (th[i] == 0.0 && n[i][1] < 0.0) th[i] = PI;
    }
}
MapPoint spheroid(const double n[8][2], double xi, double eta, double min_dist) {
    double th[8], r[8];
    theta_r(n, th, r);
    const double tt = 0.5 * ((1.0 - xi) * th[6] + (1.0 + xi) * th[4]);     // top: nodes 7, 5
    const double tb = 0.5 * ((1.0 - xi) * th[0] + (1.0 + xi) * th[2]);     // bottom: nodes 1, 3
    MapPoint p;
    p.s = 0.5 * ((1.0 + eta) * r[6] * std::sin(tt) + (1.0 - eta) * r[0] * std::sin(tb));
    p.z = 0.5 * ((1.0 + eta) * r[6] * std::cos(tt) + (1.0 - eta) * r[0] * std::cos(tb));
    if (std::fabs(p.s) < min_dist) p.s = 0.0;
    if (std::fabs(p.z) < min_dist) p.z = 0.0;
    p.dsdxi = 0.5 * ((1.0 + eta) * r[6] * 0.5 * (th[4] - th[6]) * std::cos(tt) +
                     (1.0 - eta) * r[0] * 0.5 * (th[2] - th[0]) * std::cos(tb));
    p.dzdxi = -0.5 * ((1.0 + eta) * r[6] * 0.5 * (th[4] - th[6]) * std::sin(tt) +
                      (1.0 - eta) * r[0] * 0.5 * (th[2] - th[0]) * std::sin(tb));
    p.dsdeta = 0.5 * (r[6] * std::sin(tt) - r[0] * std::sin(tb));
    p.dzdeta = 0.5 * (r[6] * std::cos(tt) - r[0] * std::cos(tb));
    return p;
}

// ---- linear: 8-node serendipity element (subpar_mapping.f90) ----------------------------------
MapPoint subpar(const double n[8][2], double xi, double eta) {
    const double xip = 1.0 + xi, xim = 1.0 - xi, etap = 1.0 + eta, etam = 1.0 - eta;
    const double xixi = xi * xi, etaeta = eta * eta;
    double shp[8], dx[8], de[8];
    shp[0] = 0.25 * xim * etam * (xim + etam - 3.0);
    shp[2] = 0.25 * xip * etam * (xip + etam - 3.0);
    shp[4] = 0.25 * xip * etap * (xip + etap - 3.0);
    shp[6] = 0.25 * xim * etap * (xim + etap - 3.0);
    shp[1] = 0.5 * etam * (1.0 - xixi);
    shp[3] = 0.5 * xip * (1.0 - etaeta);
    shp[5] = 0.5 * etap * (1.0 - xixi);
    shp[7] = 0.5 * xim * (1.0 - etaeta);
    dx[0] = -0.25 * etam * (xim + xim + etam - 3.0); de[0] = -0.25 * xim * (etam + xim + etam - 3.0);
    dx[2] = 0.25 * etam * (xip + xip + etam - 3.0);  de[2] = -0.25 * xip * (etam + xip + etam - 3.0);
    dx[4] = 0.25 * etap * (xip + xip + etap - 3.0);  de[4] = 0.25 * xip * (etap + xip + etap - 3.0);
    dx[6] = -0.25 * etap * (xim + xim + etap - 3.0); de[6] = 0.25 * xim * (etap + xim + etap - 3.0);
    dx[1] = -1.0 * xi * etam;        de[1] = -0.5 * (1.0 - xixi);
    dx[3] = 0.5 * (1.0 - etaeta);    de[3] = -1.0 * eta * xip;
    dx[5] = -1.0 * xi * etap;        de[5] = 0.5 * (1.0 - xixi);
    dx[7] = -0.5 * (1.0 - etaeta);   de[7] = -1.0 * eta * xim;
    MapPoint p{0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 8; k++) {
        p.s += shp[k] * n[k][0];
        p.z += shp[k] * n[k][1];
        p.dsdxi += n[k][0] * dx[k];
        p.dzdeta += n[k][1] * de[k];
        p.dsdeta += n[k][0] * de[k];
        p.dzdxi += n[k][1] * dx[k];
    }
    return p;
}

// ---- semi-analytic elements: one elliptic, one straight side (analytic_semi_mapping.f90) ------
void compute_ab(double &a, double &b, double s1, double z1, double s2, double z2) {
    a = std::sqrt(std::fabs((s2 * s2 * z1 * z1 - z2 * z2 * s1 * s1) / (z1 * z1 - z2 * z2)));
    b = std::sqrt(std::fabs((z1 * z1 * s2 * s2 - z2 * z2 * s1 * s1) / (s2 * s2 - s1 * s1)));
}
double ellipse_theta(double s, double z, double a, double b) {
    if (s != 0.0) return std::atan(z * a / (s * b));
    return z > 0 ? 0.5 * PI : (z < 0 ? -0.5 * PI : 0.0);
}
// straight side from node ka to node kb, elliptic side through nodes ea (xi=-1) and eb (xi=+1);
// `curved_on_top` = semino
MapPoint semi(const double n[8][2], double xi, double eta, bool curved_on_top) {
    const int ea = curved_on_top ? 6 : 0, eb = curved_on_top ? 4 : 2;      // ellipse: nodes 7,5 or 1,3
    const int la = curved_on_top ? 0 : 6, lb = curved_on_top ? 2 : 4;      // line: nodes 1,3 or 7,5
    double a, b;
    compute_ab(a, b, n[ea][0], n[ea][1], n[eb][0], n[eb][1]);
    const double tha = ellipse_theta(n[ea][0], n[ea][1], a, b), thb = ellipse_theta(n[eb][0], n[eb][1], a, b);
    const double thbar = 0.5 * (tha + thb), dth = thb - tha;
    const double arg = thbar + xi * 0.5 * dth;
    const double se = a * std::cos(arg), ze = b * std::sin(arg);
    const double dse = -a * 0.5 * dth * std::sin(arg), dze = b * 0.5 * dth * std::cos(arg);
    const double sl = 0.5 * ((1.0 + xi) * n[lb][0] + (1.0 - xi) * n[la][0]);
    const double zl = 0.5 * ((1.0 + xi) * n[lb][1] + (1.0 - xi) * n[la][1]);
    const double dsl = 0.5 * (n[lb][0] - n[la][0]), dzl = 0.5 * (n[lb][1] - n[la][1]);
    const double sbot = curved_on_top ? sl : se, zbot = curved_on_top ? zl : ze;
    const double stop = curved_on_top ? se : sl, ztop = curved_on_top ? ze : zl;
    const double dsbot = curved_on_top ? dsl : dse, dzbot = curved_on_top ? dzl : dze;
    const double dstop = curved_on_top ? dse : dsl, dztop = curved_on_top ? dze : dzl;
    const double sbar = 0.5 * (sbot + stop), ds = stop - sbot, dz = ztop - zbot;
    const double dsbar = 0.5 * (dsbot + dstop), dds = dstop - dsbot, ddz = dztop - dzbot;
    MapPoint p;
    p.s = sbar + ds * eta * 0.5;
    p.dsdxi = dsbar + 0.5 * eta * dds;
    p.dsdeta = 0.5 * ds;
    if (std::fabs(ds) > 1e-10) {
        const double intersect = (zbot * stop - ztop * sbot) / ds, slope = dz / ds;
        p.z = slope * (sbar + 0.5 * ds * eta) + intersect;
        const double dslope = (ddz * ds - dds * dz) / (ds * ds);
        const double dinter = ((dzbot * stop - dztop * sbot + zbot * dstop - ztop * dsbot) * ds -
                               dds * (zbot * stop - ztop * sbot)) / (ds * ds);
        p.dzdxi = slope * p.dsdxi + p.s * dslope + dinter;
        p.dzdeta = slope * p.dsdeta;
    } else {
        p.z = 0.5 * (zbot + ztop) + eta * (ztop - zbot) * 0.5;
        p.dzdxi = 0.0;
        p.dzdeta = 0.5 * dz;
    }
    return p;
}

}  // namespace

MapPoint map_element(int eltype, const double nodes[8][2], double xi, double eta, double min_distance_dim) {
    switch (eltype) {
    case EL_CURVED: return spheroid(nodes, xi, eta, min_distance_dim);
    case EL_LINEAR: return subpar(nodes, xi, eta);
    case EL_SEMINO: return semi(nodes, xi, eta, true);
    case EL_SEMISO: return semi(nodes, xi, eta, false);
    }
    throw std::invalid_argument("unknown element type");
}

}  // namespace axisem
