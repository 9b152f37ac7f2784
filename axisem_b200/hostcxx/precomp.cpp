#include "precomp.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <set>

#include "background_models.hpp"
#include "mapping.hpp"

namespace axisem {
namespace {

constexpr int NP = 5, NPT = 25;
const double PI = 3.14159265358979323846;

Array make(Array::Type t, std::vector<uint64_t> dims, const void *src) {
    Array a;
    a.type = t;
    a.dims = std::move(dims);
    const size_t nb = a.count() * (t == Array::F64 ? 8 : 4);
    a.bytes.resize(nb);
    if (nb) std::memcpy(a.bytes.data(), src, nb);
    return a;
}
Array f32_of(const std::vector<double> &v, std::vector<uint64_t> dims) {
    std::vector<float> f(v.size());
    for (size_t k = 0; k < v.size(); k++) f[k] = (float)v[k];
    return make(Array::F32, std::move(dims), f.data());
}
Array i32_of(const std::vector<int32_t> &v, std::vector<uint64_t> dims) { return make(Array::I32, std::move(dims), v.data()); }
Array scalar_i(int32_t v) { return make(Array::I32, {}, &v); }
Array scalar_d(double v) { return make(Array::F64, {}, &v); }

// ---- element geometry of one domain of one rank ----------------------------------------------
struct Geometry {
    int nel = 0;
    std::vector<int32_t> axis, iel_glob;            // per element
    std::vector<double> s, z, r, sin_t, cos_t;      // per point (e*25 + 5*j + i)
    std::vector<double> dsdxi, dzdxi, dsdeta, dzdeta, jac, W, W2, massmat_k, kwts2;
    std::vector<double> s_gll;                      // s at the GLL abscissa (the reference's anelastic quirk)
    std::vector<double> r_mid;                      // per element: radius of the mid-point
};

struct Spectral { const double *eta, *wt, *xi_k, *wt_axial_k; };

Geometry element_geometry(const Modules &m, bool fluid, const Spectral &sp) {
    Geometry g;
    const std::string dom = fluid ? "fluid" : "solid";
    g.nel = m.int_of("data_mesh%nel_" + dom);
    const int nelem = m.int_of("data_mesh%nelem");
    const int npoin = m.int_of("data_mesh%npoin");
    const int32_t *iel_of = m.i(fluid ? "data_mesh%ielfluid" : "data_mesh%ielsolid");
    const int32_t *axis = m.i("data_mesh%axis_" + dom);
    const int32_t *lnods = m.i("data_mesh%lnods");      // Fortran (nelem, 8)
    const double *crd = m.d("data_mesh%crd_nodes");     // Fortran (npoin, 2)
    const int32_t *eltype = m.i("data_mesh%eltype");
    const double min_dist = 1e-3;                        // metres: masks round-off on the axis (get_mesh.f90:188)
    const size_t n = (size_t)NPT * g.nel;
    for (auto *v : {&g.s, &g.z, &g.r, &g.sin_t, &g.cos_t, &g.dsdxi, &g.dzdxi, &g.dsdeta, &g.dzdeta, &g.jac, &g.W, &g.W2,
                    &g.massmat_k, &g.kwts2, &g.s_gll})
        v->assign(n, 0.0);
    g.axis.assign(g.nel, 0);
    g.iel_glob.assign(g.nel, 0);
    g.r_mid.assign(g.nel, 0.0);
    for (int e = 0; e < g.nel; e++) {
        const int ig = iel_of[e] - 1;
        g.iel_glob[e] = ig;
        g.axis[e] = axis[e] != 0;
        double nodes[8][2];
        for (int k = 0; k < 8; k++) {
            const int nd = lnods[(size_t)k * nelem + ig] - 1;
            nodes[k][0] = crd[nd];
            nodes[k][1] = crd[(size_t)npoin + nd];
        }
        const bool ax = g.axis[e];
        {
            const MapPoint c = map_element(eltype[ig], nodes, 0.0, 0.0, min_dist);
            g.r_mid[e] = std::sqrt(c.s * c.s + c.z * c.z);
        }
        for (int j = 0; j < NP; j++)
            for (int i = 0; i < NP; i++) {
                const size_t p = (size_t)NPT * e + NP * j + i;
                const double xi = ax ? sp.xi_k[i] : sp.eta[i];
                MapPoint mp = map_element(eltype[ig], nodes, xi, sp.eta[j], min_dist);
                if (ax && i == 0) mp.s = 0.0;
                g.s[p] = mp.s; g.z[p] = mp.z;
                g.dsdxi[p] = mp.dsdxi; g.dzdxi[p] = mp.dzdxi; g.dsdeta[p] = mp.dsdeta; g.dzdeta[p] = mp.dzdeta;
                g.jac[p] = mp.jacobian();
                const double r = std::sqrt(mp.s * mp.s + mp.z * mp.z);
                g.r[p] = r;
                g.sin_t[p] = r > 0 ? mp.s / r : 0.0;
                g.cos_t[p] = r > 0 ? mp.z / r : 1.0;
                const double wxi = ax ? sp.wt_axial_k[i] : sp.wt[i];
                const double ww = sp.wt[j] * wxi;
                const double opx = 1.0 + xi;
                // s / (1 + xi), on the axis its limit ds/dxi (analytic_mapping.f90:71-93)
                const double sop = (ax && i == 0) ? mp.dsdxi : mp.s / opx;
                g.W[p] = (ax ? sop : mp.s) * ww;
                g.W2[p] = ww * (ax ? (opx > 0 ? 1.0 / opx : 0.0) : 1.0);
                g.massmat_k[p] = g.jac[p] * g.W[p];
                // massmat_kwts2 (def_precomp_terms.f90:673-707)
                if (!ax) g.kwts2[p] = g.jac[p] / mp.s * ww;
                else if (i == 0) g.kwts2[p] = g.jac[p] / sop * ww;
                else g.kwts2[p] = g.jac[p] / (mp.s * opx) * ww;
                g.s_gll[p] = ax ? map_element(eltype[ig], nodes, sp.eta[i], sp.eta[j], min_dist).s : mp.s;
            }
    }
    return g;
}

// ---- background model (get_model.F90:155-186) ------------------------------------------------
struct Material {
    std::vector<double> rho, lam, mu, xi, phi, eta, vp;   // per point
    std::vector<double> qmu, qka;                          // per element
};
Material material(const std::string &model, const Geometry &g) {
    Material M;
    const size_t n = (size_t)NPT * g.nel;
    for (auto *v : {&M.rho, &M.lam, &M.mu, &M.xi, &M.phi, &M.eta, &M.vp}) v->assign(n, 0.0);
    M.qmu.assign(g.nel, 0.0);
    M.qka.assign(g.nel, 0.0);
    for (int e = 0; e < g.nel; e++) {
        const int idom = model_domain_of(model, g.r_mid[e], false);
        for (int q = 0; q < NPT; q++) {
            const size_t p = (size_t)NPT * e + q;
            const ModelValues v = model_evaluate(model, g.r[p], idom);
            M.rho[p] = v.rho;
            M.lam[p] = v.rho * (v.vph * v.vph - 2.0 * v.vsh * v.vsh);
            M.mu[p] = v.rho * v.vsh * v.vsh;
            M.xi[p] = v.vsv > 1e-10 * v.vph ? v.vsh * v.vsh / (v.vsv * v.vsv) : 1.0;
            M.phi[p] = v.vpv * v.vpv / (v.vph * v.vph);
            M.eta[p] = v.eta;
            M.vp[p] = std::max(v.vph, v.vpv);
            if (q == NP * (NP / 2 - 1) + NP / 2 - 1) {      // Q_mu_1d at (npol/2 - 1, npol/2 - 1), get_model.F90:236-242
                M.qmu[e] = v.qmu;
                M.qka[e] = v.qkappa;
            }
        }
    }
    return M;
}

// def_precomp_terms.f90:2284-2332 with the fast axis s = (sin th, 0, cos th) (radial TI)
double c_ijkl(double lam, double mu, double xi, double phi, double eta, double sin_fa, double cos_fa, int i, int j,
              int k, int l) {
    auto d = [](int a, int b) { return a == b ? 1.0 : 0.0; };
    const double s[4] = {0.0, sin_fa, 0.0, cos_fa};
    double c = 0.0;
    c = c + lam * d(i, j) * d(k, l);
    c = c + mu * (d(i, k) * d(j, l) + d(i, l) * d(j, k));
    c = c + ((eta - 1.0) * lam + 2.0 * eta * mu * (1.0 - 1.0 / xi)) * (d(i, j) * s[k] * s[l] + d(k, l) * s[i] * s[j]);
    c = c + mu * (1.0 / xi - 1.0) *
                (d(i, k) * s[j] * s[l] + d(i, l) * s[j] * s[k] + d(j, k) * s[i] * s[l] + d(j, l) * s[i] * s[k]);
    c = c + ((1.0 - 2.0 * eta + phi) * (lam + 2.0 * mu) + (4.0 * eta - 4.0) * mu / xi) * (s[i] * s[j] * s[k] * s[l]);
    return c;
}

// ---- solid stiffness planes (def_precomp_terms.f90:1166-2332) --------------------------------
using Planes = std::map<std::string, std::vector<double>>;

void solid_stiffness_terms(const std::string &src, const Geometry &g, const Material &M, const std::vector<double> &lam,
                           const std::vector<double> &mu, bool anel, const Spectral &sp, Planes &out, Planes &out0,
                           Planes &out_cg) {
    const size_t n = (size_t)NPT * g.nel;
    auto plane = [&](const char *name) -> std::vector<double> & { out[name].assign(n, 0.0); return out[name]; };
    auto vec0 = [&](const char *name) -> std::vector<double> & { out0[name].assign((size_t)NP * g.nel, 0.0); return out0[name]; };
    const bool mono = src == "monopole", di = src == "dipole", quad = src == "quadpole";
    std::vector<const char *> names = {"M11s", "M21s", "M41s", "M12s", "M22s", "M32s", "M42s", "M11z", "M21z", "M41z",
                                       "M_1", "M_2", "M_3", "M_4", "M_w1"};
    if (di) for (const char *s : {"M13s", "M33s", "M43s", "M_5", "M_6", "M_7", "M_8", "M_w2", "M_w3"}) names.push_back(s);
    if (quad) for (const char *s : {"M1phi", "M2phi", "M4phi", "M_5", "M_6", "M_7", "M_8", "M_w2", "M_w3", "M_w4", "M_w5"}) names.push_back(s);
    for (const char *s : names) plane(s);
    const int nw0 = mono ? 3 : (di ? 10 : 6);
    for (int k = 1; k <= nw0; k++) vec0(("M0_w" + std::to_string(k)).c_str());
    if (anel) {
        for (const char *s : {"Y", "V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi"}) plane(s);
        for (const char *s : {"Y0", "V0_s_eta", "V0_s_xi", "V0_z_eta", "V0_z_xi"}) vec0(s);
    }
    for (int e = 0; e < g.nel; e++) {
        const bool ax = g.axis[e];
        for (int j = 0; j < NP; j++)
            for (int i = 0; i < NP; i++) {
                const size_t p = (size_t)NPT * e + NP * j + i;
                const double ij = 1.0 / g.jac[p], W = g.W[p], W2 = g.W2[p];
                const double dsdxi = g.dsdxi[p], dzdxi = g.dzdxi[p], dsdeta = g.dsdeta[p], dzdeta = g.dzdeta[p];
                // analytic_mapping.f90:114-570 (the s factor and quadrature weights are inside W)
                const double alpha = -ij * dsdxi * dsdeta * W, beta = ij * dsdxi * dsdxi * W, gamma = ij * dsdeta * dsdeta * W;
                const double delta = -ij * dzdxi * dzdeta * W, epsil = ij * dzdxi * dzdxi * W, zeta = ij * dzdeta * dzdeta * W;
                const double Ms_ze_sx = ij * dsdxi * dzdeta * W, Ms_ze_se = -ij * dsdeta * dzdeta * W;
                const double Ms_zx_se = ij * dsdeta * dzdxi * W, Ms_zx_sx = -ij * dsdxi * dzdxi * W;
                const double M_s_xi = dsdxi * W2, M_z_xi = -dzdxi * W2, M_z_eta = dzdeta * W2, M_s_eta = -dsdeta * W2;
                const double kw2 = g.kwts2[p];
                auto C = [&](int a, int b, int c, int d) {
                    return c_ijkl(lam[p], mu[p], M.xi[p], M.phi[p], M.eta[p], g.sin_t[p], g.cos_t[p], a, b, c, d);
                };
                const double C11 = C(1, 1, 1, 1), C12 = C(1, 1, 2, 2), C13 = C(1, 1, 3, 3), C15 = C(1, 1, 3, 1);
                const double C22 = C(2, 2, 2, 2), C23 = C(2, 2, 3, 3), C25 = C(2, 2, 3, 1);
                const double C33 = C(3, 3, 3, 3), C35 = C(3, 3, 3, 1);
                const double C44 = C(2, 3, 2, 3), C46 = C(2, 3, 1, 2), C55 = C(3, 1, 3, 1), C66 = C(1, 2, 1, 2);
                const bool i0 = ax && i == 0;
                auto set = [&](const char *nm, double v, bool zero_on_axis = false) { out[nm][p] = (zero_on_axis && i0) ? 0.0 : v; };
                if (mono || quad) {
                    set("M11s", C11 * delta + C15 * Ms_ze_sx + C15 * Ms_zx_se + C55 * alpha);
                    set("M21s", C11 * zeta + C15 * 2.0 * Ms_ze_se + C55 * gamma);
                    set("M41s", C11 * epsil + C15 * 2.0 * Ms_zx_sx + C55 * beta);
                    set("M12s", C15 * delta + C13 * Ms_ze_sx + C55 * Ms_zx_se + C35 * alpha);
                    set("M22s", C15 * zeta + (C13 + C55) * Ms_ze_se + C35 * gamma);
                    set("M32s", C15 * delta + C13 * Ms_zx_se + C55 * Ms_ze_sx + C35 * alpha);
                    set("M42s", C15 * epsil + (C13 + C55) * Ms_zx_sx + C35 * beta);
                    set("M11z", C55 * delta + C35 * Ms_ze_sx + C35 * Ms_zx_se + C33 * alpha);
                    set("M21z", C55 * zeta + C35 * 2.0 * Ms_ze_se + C33 * gamma);
                    set("M41z", C55 * epsil + C35 * 2.0 * Ms_zx_sx + C33 * beta);
                    set("M_1", C12 * M_z_eta + C25 * M_s_eta, quad);
                    set("M_2", C12 * M_z_xi + C25 * M_s_xi, quad);
                    set("M_3", C23 * M_s_eta + C25 * M_z_eta, quad);
                    set("M_4", C23 * M_s_xi + C25 * M_z_xi, quad);
                }
                if (mono) {
                    set("M_w1", C22 * kw2, true);
                } else if (di) {
                    const double sum_ms = Ms_ze_sx + Ms_zx_se;
                    set("M11s", (C11 + C66) * delta + (C15 + C46) * sum_ms + (C55 + C44) * alpha);
                    set("M21s", (C11 + C66) * zeta + (C15 + C46) * 2.0 * Ms_ze_se + (C55 + C44) * gamma);
                    set("M41s", (C11 + C66) * epsil + (C15 + C46) * 2.0 * Ms_zx_sx + (C55 + C44) * beta);
                    set("M12s", (C11 - C66) * delta + (C15 - C46) * sum_ms + (C55 - C44) * alpha);
                    set("M22s", (C11 - C66) * zeta + (C15 - C46) * 2.0 * Ms_ze_se + (C55 - C44) * gamma);
                    set("M42s", (C11 - C66) * epsil + (C15 - C46) * 2.0 * Ms_zx_sx + (C55 - C44) * beta);
                    set("M13s", C15 * delta + C13 * Ms_ze_sx + C55 * Ms_zx_se + C35 * alpha);
                    set("M32s", C15 * zeta + (C13 + C55) * Ms_ze_se + C35 * gamma);
                    set("M33s", C15 * delta + C13 * Ms_zx_se + C55 * Ms_ze_sx + C35 * alpha);
                    set("M43s", C15 * epsil + (C13 + C55) * Ms_zx_sx + C35 * beta);
                    set("M11z", C55 * delta + C35 * sum_ms + C33 * alpha);
                    set("M21z", C55 * zeta + C35 * 2.0 * Ms_ze_se + C33 * gamma);
                    set("M41z", C55 * epsil + C35 * 2.0 * Ms_zx_sx + C33 * beta);
                    set("M_1", (C12 + C66) * 2.0 * M_z_eta + (C25 + C46) * 2.0 * M_s_eta, true);
                    set("M_2", (C12 + C66) * 2.0 * M_z_xi + (C25 + C46) * 2.0 * M_s_xi, true);
                    set("M_3", C46 * M_z_eta + C44 * M_s_eta, true);
                    set("M_4", C46 * M_z_xi + C44 * M_s_xi, true);
                    set("M_5", (C12 - C66) * 2.0 * M_z_eta + (C25 - C46) * 2.0 * M_s_eta, true);
                    set("M_6", (C12 - C66) * 2.0 * M_z_xi + (C25 - C46) * 2.0 * M_s_xi, true);
                    set("M_7", C25 * 2.0 * M_z_eta + C23 * 2.0 * M_s_eta, true);
                    set("M_8", C25 * 2.0 * M_z_xi + C23 * 2.0 * M_s_xi, true);
                    set("M_w1", 4.0 * (C22 + C66) * kw2, true);
                    set("M_w2", 2.0 * C46 * kw2);
                    set("M_w3", C44 * kw2, true);
                } else {
                    set("M1phi", C66 * delta + C46 * Ms_ze_sx + C46 * Ms_zx_se + C44 * alpha);
                    set("M2phi", C66 * zeta + C46 * 2.0 * Ms_ze_se + C44 * gamma);
                    set("M4phi", C66 * epsil + C46 * 2.0 * Ms_zx_sx + C44 * beta);
                    set("M_5", C66 * M_z_eta + C46 * M_s_eta, true);
                    set("M_6", C66 * M_z_xi + C46 * M_s_xi, true);
                    set("M_7", C44 * M_s_eta + C46 * M_z_eta, true);
                    set("M_8", C44 * M_s_xi + C46 * M_z_xi, true);
                    set("M_w1", (C22 + 4.0 * C66) * kw2, true);
                    set("M_w2", -2.0 * (C22 + C66) * kw2, true);
                    set("M_w3", 2.0 * C46 * kw2, true);
                    set("M_w4", (4.0 * C22 + C66) * kw2, true);
                    set("M_w5", 4.0 * C44 * kw2, true);
                }
                if (anel) {
                    // def_precomp_terms.f90:1400-1416, 1481-1526; in axial elements the reference takes
                    // s at the GLL abscissa eta(ipol) instead of xi_k(ipol) (:1486)
                    const double s_use = g.s_gll[p];
                    out["Y"][p] = i0 ? 0.0 : W2 * g.jac[p];
                    out["V_s_eta"][p] = i0 ? 0.0 : s_use * M_s_eta;
                    out["V_s_xi"][p] = i0 ? 0.0 : s_use * M_s_xi;
                    out["V_z_eta"][p] = i0 ? 0.0 : s_use * M_z_eta;
                    out["V_z_xi"][p] = i0 ? 0.0 : s_use * M_z_xi;
                }
                if (ax && i == 0) {
                    // axial vectors live at ipol = 0 (def_precomp_terms.f90:1316-1329)
                    const size_t a = (size_t)NP * e + j;
                    const double ndf = kw2, w0 = sp.wt_axial_k[0] * sp.wt[j];
                    auto v0 = [&](const char *nm, double v) { out0[nm][a] = v; };
                    if (mono) {
                        v0("M0_w1", (2.0 * C12 + C22) * ndf);
                        v0("M0_w2", C25 * ndf);
                        v0("M0_w3", C23 * dsdxi * w0 - C25 * dzdxi * w0);
                    } else if (di) {
                        v0("M0_w1", (C12 + C66) * 2.0 * ndf);
                        v0("M0_w2", -(C12 + C66) * 2.0 * dzdxi * w0);
                        v0("M0_w3", C46 * ndf);
                        v0("M0_w4", -C46 * dzdxi * w0);
                        v0("M0_w6", (C25 + C46) * 2.0 * dsdxi * w0);
                        v0("M0_w7", C44 * ndf);
                        v0("M0_w8", C44 * dsdxi * w0);
                        v0("M0_w9", (C12 + C22) * 4.0 * ndf);
                        v0("M0_w10", (2.0 * C25 + C46) * ndf);
                    } else {
                        v0("M0_w1", (2.0 * C12 + C22 + 4.0 * C66) * ndf);
                        v0("M0_w2", -2.0 * (C12 + C22) * ndf);
                        v0("M0_w3", (C25 + 4.0 * C46) * ndf);
                        v0("M0_w4", (4.0 * C22 - C66) * ndf);
                        v0("M0_w5", -2.0 * C25 * ndf);
                        v0("M0_w6", 4.0 * C44 * ndf);
                    }
                    if (anel) {
                        v0("Y0", w0 * g.jac[p]);
                        v0("V0_s_xi", w0 * dsdxi * dsdxi);
                        v0("V0_z_eta", w0 * dsdxi * dzdeta);
                        v0("V0_z_xi", w0 * dsdxi * (-dzdxi));
                    }
                }
            }
    }
    if (anel)
        for (const char *nm : {"Y", "V_s_eta", "V_s_xi", "V_z_eta", "V_z_xi"}) {
            std::vector<double> &c = out_cg[std::string(nm) + "_cg4"];
            c.assign((size_t)4 * g.nel, 0.0);
            const std::vector<double> &a = out[nm];
            for (int e = 0; e < g.nel; e++) {
                // A(1,1), A(1,3), A(3,1), A(3,3) in (ipol, jpol) (def_precomp_terms.f90:1581-1604)
                const size_t b = (size_t)NPT * e;
                c[4 * e + 0] = a[b + NP * 1 + 1]; c[4 * e + 1] = a[b + NP * 3 + 1];
                c[4 * e + 2] = a[b + NP * 1 + 3]; c[4 * e + 3] = a[b + NP * 3 + 3];
            }
        }
}

std::vector<double> cg4(const std::vector<double> &a, int nel) {
    std::vector<double> c((size_t)4 * nel);
    for (int e = 0; e < nel; e++) {
        const size_t b = (size_t)NPT * e;
        c[4 * e + 0] = a[b + NP * 1 + 1]; c[4 * e + 1] = a[b + NP * 3 + 1];
        c[4 * e + 2] = a[b + NP * 1 + 3]; c[4 * e + 3] = a[b + NP * 3 + 3];
    }
    return c;
}

// attenuation.f90:1139-1155
std::vector<double> fast_correct(const std::vector<double> &y) {
    std::vector<double> dy(y.size()), yp(y.size());
    dy[0] = 1.0 + 0.5 * y[0];
    for (size_t k = 1; k < y.size(); k++) dy[k] = dy[k - 1] + (dy[k - 1] - 0.5) * y[k - 1] + 0.5 * y[k];
    for (size_t k = 0; k < y.size(); k++) yp[k] = y[k] * dy[k];
    return yp;
}

// direct stiffness summation of a per-point field over the rank's global numbers, plus the halo
// partners' sums (what pdistsum_* does to the mass matrix, def_precomp_terms.f90:756, :820)
void assemble(std::vector<Modules> &ranks, const std::string &dom, std::vector<std::vector<double>> &val) {
    const size_t nr = ranks.size();
    std::vector<std::vector<double>> glob(nr);
    for (size_t r = 0; r < nr; r++) {
        const int nglob = ranks[r].int_of("data_mesh%nglob_" + dom);
        const int32_t *ig = ranks[r].i("data_mesh%igloc_" + dom);
        glob[r].assign((size_t)nglob, 0.0);
        for (size_t p = 0; p < val[r].size(); p++) glob[r][ig[p] - 1] += val[r][p];
    }
    std::vector<std::vector<double>> add(nr);
    for (size_t r = 0; r < nr; r++) add[r].assign(glob[r].size(), 0.0);
    for (size_t r = 0; r < nr; r++) {
        const int nmsg = ranks[r].int_of("data_comm%sizerecv_" + dom, 0);
        if (nmsg <= 0) continue;
        const int32_t *peer = ranks[r].i("data_comm%listrecv_" + dom), *size = ranks[r].i("data_comm%sizemsgrecv_" + dom);
        const Array &gl = ranks[r].at("data_comm%glocal_index_msg_recv_" + dom);
        const int maxmsg = (int)gl.dims[1];
        for (int m = 0; m < nmsg; m++) {
            // the peer's list towards me holds the same points in the same order (get_mesh.f90:303-310)
            size_t pr = nr;
            for (size_t q = 0; q < nr; q++) if (ranks[q].int_of("data_proc%mynum") == peer[m]) pr = q;
            if (pr == nr) throw SolverError("precompute: halo peer missing among the ranks given");
            const int nm2 = ranks[pr].int_of("data_comm%sizerecv_" + dom, 0);
            const int32_t *peer2 = ranks[pr].i("data_comm%listrecv_" + dom);
            const Array &gl2 = ranks[pr].at("data_comm%glocal_index_msg_recv_" + dom);
            int mm = -1;
            for (int q = 0; q < nm2; q++) if (peer2[q] == ranks[r].int_of("data_proc%mynum")) mm = q;
            if (mm < 0) throw SolverError("precompute: halo lists inconsistent between ranks");
            for (int ip = 0; ip < size[m]; ip++)
                add[r][gl.i32()[(size_t)m * maxmsg + ip] - 1] += glob[pr][gl2.i32()[(size_t)mm * gl2.dims[1] + ip] - 1];
        }
    }
    for (size_t r = 0; r < nr; r++) {
        const int32_t *ig = ranks[r].i("data_mesh%igloc_" + dom);
        for (size_t p = 0; p < val[r].size(); p++) val[r][p] = glob[r][ig[p] - 1] + add[r][ig[p] - 1];
    }
}

// shift_fact of the reference in seconds
double stf_shift(const PrecompOptions &o, double deltat) {
    return o.shift_seconds >= 0.0 ? o.shift_seconds : std::ceil(o.shift_fact * o.t_0 / deltat) * deltat;
}

// the reference's own error function (Numerical Recipes erfc, source.f90:662-692; the coefficients
// are default-real literals)
double erf_nr(double x) {
    static const float c[10] = {-1.26551223f, 1.00002368f, 0.37409196f, 0.09678418f, -0.18628806f,
                                0.27886807f, -1.13520398f, 1.48851587f, -0.82215223f, 0.17087277f};
    const double z = std::fabs(x), t = 1.0 / (1.0 + 0.5 * z);
    double poly = (double)c[9];
    for (int k = 8; k >= 0; k--) poly = t * poly + (double)c[k];
    double erfcc = t * std::exp(-z * z + poly);
    if (x < 0.0) erfcc = 2.0 - erfcc;
    return 1.0 - erfcc;
}

// the smooth source time functions: gauss, gauss_d, gauss_dd, errorf (source.f90:587-660)
double stf_at(const PrecompOptions &o, double t, double deltat) {
    const double shift = stf_shift(o, deltat), a = o.decay / o.t_0, x = a * (t - shift);
    if (o.stf_type == "gauss_0") return std::exp(-x * x) * o.magnitude * a / std::sqrt(PI);
    if (o.stf_type == "gauss_1") return -2.0 * a * a * (t - shift) * std::exp(-x * x) / (a * std::sqrt(2.0) * std::exp(-0.5)) * o.magnitude;
    if (o.stf_type == "gauss_2")
        return a * a * (2.0 * a * a * (t - shift) * (t - shift) - 1.0) * std::exp(-x * x) / (2.0 * a * a * std::exp(-1.5)) * o.magnitude;
    if (o.stf_type == "errorf") return (erf_nr(x) * 0.5 + 0.5) * o.magnitude;
    throw SolverError("source time function non existant: " + o.stf_type);
}

// delta_src (source.f90:696-814): the discrete Dirac of `discrete_choice`, normalised to unit integral,
// times the magnitude; quheavi: its running integral
std::vector<float> delta_src(const PrecompOptions &o, int niter, double deltat) {
    const double gpi = 3.1415926535898;                       // global_parameters.f90
    const double a = (double)(float)o.t_0;                    // discrete_dirac_halfwidth (realkind)
    const double shift = (double)(float)stf_shift(o, deltat); // shift_fact_discrete_dirac (realkind)
    const std::string &c = o.discrete_choice;
    std::vector<double> signal(niter, 0.0);
    for (int i = 1; i <= niter; i++) {
        double t = (double)i * deltat, v;
        if (c == "cauchy") v = 1.0 / a * std::exp(-std::fabs((t - shift) / a));
        else if (c == "caulor") v = 1.0 / gpi * a / (a * a + (t - shift) * (t - shift));
        else if (c == "sincfc") {
            if (t == shift) t = 0.00001 + shift;
            v = 1.0 / (a * gpi) * std::sin((-shift + t) / a) / ((-shift + t) / a);
        } else if (c == "gaussi") v = 1.0 / (a * std::sqrt(gpi)) * std::exp(-((t - shift) / a) * ((t - shift) / a));
        else if (c == "triang") v = std::fabs(t - shift) <= a / 2.0 ? 2.0 / a - 4.0 / (a * a) * std::fabs(t - shift) : 0.0;
        else if (c == "1dirac") v = i == (int)(shift / deltat) ? 1.0 : 0.0;
        else throw SolverError("do not know discrete Dirac " + c);
        signal[i - 1] = v;
    }
    double sum = 0.0;
    for (double v : signal) sum += v;
    const double integral = sum * deltat;
    std::vector<float> stf(niter);
    for (int i = 0; i < niter; i++) stf[i] = (float)((double)(float)(signal[i] / integral) * o.magnitude);
    if (o.stf_type == "quheavi") {
        double acc = 0.0;
        for (int i = 0; i < niter; i++) { acc += (double)stf[i] * deltat; stf[i] = (float)acc; }
    }
    return stf;
}

// compute_stf (source.f90:144-202)
std::vector<float> compute_stf(const PrecompOptions &o, int niter, double deltat) {
    if (o.stf_type == "dirac_0" || o.stf_type == "dirac_1" || o.stf_type == "quheavi") return delta_src(o, niter, deltat);
    std::vector<float> stf(niter);
    for (int k = 0; k < niter; k++) stf[k] = (float)stf_at(o, (double)(float)((k + 1) * deltat), deltat);
    return stf;
}

int stf_code(const std::string &s) {       // include/axisem_b200.h
    static const std::map<std::string, int> codes = {{"gauss_0", 0}, {"gauss_1", 1}, {"gauss_2", 2}, {"errorf", 3},
                                                     {"dirac_0", 4}, {"dirac_1", 4}, {"quheavi", 5}};
    const auto it = codes.find(s);
    if (it == codes.end()) throw SolverError("source time function non existant: " + s);
    return it->second;
}

const std::map<std::string, std::string> SRC_POLE = {
    {"explosion", "monopole"}, {"mrr", "monopole"}, {"mtt_p_mpp", "monopole"}, {"vertforce", "monopole"},
    {"mtr", "dipole"}, {"mpr", "dipole"}, {"thetaforce", "dipole"}, {"phiforce", "dipole"},
    {"mtp", "quadpole"}, {"mtt_m_mpp", "quadpole"}};

}  // namespace

// ================================================================================================
void precompute(std::vector<Modules> &ranks, const PrecompOptions &opt) {
    if (ranks.empty()) return;
    const auto itp = SRC_POLE.find(opt.src_type2);
    if (itp == SRC_POLE.end()) throw SolverError("unknown source type " + opt.src_type2);
    const std::string src = itp->second;
    const int src_order = src == "monopole" ? 0 : (src == "dipole" ? 1 : 2);
    const size_t nr = ranks.size();
    std::vector<Geometry> gs(nr), gf(nr);
    std::vector<Material> ms(nr), mf(nr);
    std::vector<std::vector<double>> mass_s(nr), mass_f(nr);
    std::string model = opt.model;
    double deltat = opt.deltat;
    for (size_t r = 0; r < nr; r++) {
        Modules &m = ranks[r];
        if (m.int_of("data_mesh%npol") != 4) throw SolverError("precompute: npol must be 4");
        if (model.empty()) {
            const Array &b = m.at("data_mesh%bkgrdmodel");
            for (size_t k = 0; k < b.count(); k++) model.push_back((char)b.i32()[k]);
        }
        if (opt.attenuation && !model_is_anelastic(model))          // get_mesh.f90:142-150
            throw SolverError("viscoelastic attenuation set, but backgroundmodel " + model + " is elastic only.");
        if (deltat <= 0) deltat = m.real_of("data_time%deltat") * (opt.time_scheme == "newmark2" ? 1.0 : 1.5);
        const Spectral sp{m.d("data_spec%eta"), m.d("data_spec%wt"), m.d("data_spec%xi_k"), m.d("data_spec%wt_axial_k")};
        gs[r] = element_geometry(m, false, sp);
        gf[r] = element_geometry(m, true, sp);
        ms[r] = material(model, gs[r]);
        mf[r] = material(model, gf[r]);
        mass_s[r].resize(gs[r].s.size());
        for (size_t p = 0; p < mass_s[r].size(); p++) mass_s[r][p] = (double)(float)(ms[r].rho[p] * gs[r].massmat_k[p]);
        mass_f[r].resize(gf[r].s.size());
        for (size_t p = 0; p < mass_f[r].size(); p++) mass_f[r][p] = (double)(float)(gf[r].massmat_k[p] / mf[r].lam[p]);
    }
    // mass matrices (before the attenuation changes the moduli, as in the reference)
    std::vector<std::vector<double>> um_s = mass_s, um_f = mass_f;
    assemble(ranks, "solid", mass_s);
    assemble(ranks, "fluid", mass_f);

    const int n_sls = (int)opt.att.w_j.size();
    for (size_t r = 0; r < nr; r++) {
        Modules &m = ranks[r];
        const Spectral sp{m.d("data_spec%eta"), m.d("data_spec%wt"), m.d("data_spec%xi_k"), m.d("data_spec%wt_axial_k")};
        const Geometry &g = gs[r], &f = gf[r];
        const int nel_s = g.nel, nel_f = f.nel;
        const double router = m.real_of("data_mesh%router");
        auto put25 = [&](const std::string &name, const std::vector<double> &v, int nel) { m.put(name, f32_of(v, {(uint64_t)nel, NP, NP})); };
        m.put("data_source%src_order", scalar_i(src_order));
        // ---- mass ---------------------------------------------------------------------------
        {
            std::vector<double> inv(mass_s[r].size());
            for (size_t p = 0; p < inv.size(); p++) inv[p] = (src == "dipole" ? 0.5 : 1.0) / mass_s[r][p];   // :773
            put25("data_matr%inv_mass_rho", inv, nel_s);
            if (nel_f) {
                std::vector<double> invf(mass_f[r].size());
                for (size_t p = 0; p < invf.size(); p++) invf[p] = 1.0 / mass_f[r][p];
                put25("data_matr%inv_mass_fluid", invf, nel_f);
            }
            if (opt.dump_energy) {
                std::vector<double> u = um_s[r];
                if (src == "dipole") for (double &x : u) x = (double)(float)x * 2.0;
                put25("data_matr%unassem_mass_rho_solid", u, nel_s);
                if (nel_f) put25("data_matr%unassem_mass_lam_fluid", um_f[r], nel_f);
            }
        }
        // ---- pointwise derivative planes (def_precomp_terms.f90:178-225) ------------------------
        auto pointwise = [&](const Geometry &G, const std::string &suffix, int nel) {
            const size_t n = G.s.size();
            std::vector<double> dse(n), dze(n), dsx(n), dzx(n), inv_s(n);
            for (int e = 0; e < nel; e++)
                for (int q = 0; q < NPT; q++) {
                    const size_t p = (size_t)NPT * e + q;
                    dse[p] = -G.dsdeta[p] / G.jac[p]; dze[p] = G.dzdeta[p] / G.jac[p];
                    dsx[p] = G.dsdxi[p] / G.jac[p];   dzx[p] = -G.dzdxi[p] / G.jac[p];
                    inv_s[p] = (G.s[p] != 0.0 && !(G.axis[e] && q % NP == 0)) ? 1.0 / G.s[p] : 1.0;
                }
            put25("data_pointwise%DsDeta_over_J_" + suffix, dse, nel);
            put25("data_pointwise%DzDeta_over_J_" + suffix, dze, nel);
            put25("data_pointwise%DsDxi_over_J_" + suffix, dsx, nel);
            put25("data_pointwise%DzDxi_over_J_" + suffix, dzx, nel);
            put25("data_pointwise%inv_s_" + std::string(suffix == "sol" ? "solid" : "fluid"), inv_s, nel);
        };
        pointwise(g, "sol", nel_s);
        if (nel_f) pointwise(f, "flu", nel_f);
        // ---- attenuation (attenuation.f90:882-1063) -----------------------------------------------
        std::vector<double> lam_u = ms[r].lam, mu_u = ms[r].mu;
        m.put("attenuation%anel_true", scalar_i(opt.attenuation));
        if (opt.attenuation) {
            const AttenuationOptions &A = opt.att;
            std::vector<double> exp_w(n_sls), ts_t(n_sls), ts_tm1(n_sls);
            for (int k = 0; k < n_sls; k++) {
                exp_w[k] = std::exp(-A.w_j[k] * deltat);
                ts_tm1[k] = (1.0 - exp_w[k]) / (A.w_j[k] * deltat) - exp_w[k];
                ts_t[k] = (exp_w[k] - 1.0) / (A.w_j[k] * deltat) + 1.0;
            }
            const double w_0 = A.w_0 * 2 * PI, w_1 = std::sqrt(A.f_min * A.f_max) * 2 * PI;
            const size_t n = g.s.size();
            std::vector<double> dmu(n), dka(n), wcg(n, 0.0);
            for (int e = 0; e < nel_s; e++) {
                const size_t b = (size_t)NPT * e;
                auto G = [&](int i, int j) { return g.massmat_k[b + NP * j + i]; };
                // coarse-grained weights (attenuation.f90:940-996)
                wcg[b + NP * 1 + 1] = (G(0, 0) + G(0, 1) + G(1, 0) + G(1, 1) + 0.5 * (G(0, 2) + G(1, 2) + G(2, 0) + G(2, 1)) + 0.25 * G(2, 2)) / G(1, 1);
                wcg[b + NP * 1 + 3] = (G(3, 0) + G(3, 1) + G(4, 0) + G(4, 1) + 0.5 * (G(2, 0) + G(2, 1) + G(3, 2) + G(4, 2)) + 0.25 * G(2, 2)) / G(3, 1);
                wcg[b + NP * 3 + 1] = (G(0, 3) + G(0, 4) + G(1, 3) + G(1, 4) + 0.5 * (G(0, 2) + G(1, 2) + G(2, 3) + G(2, 4)) + 0.25 * G(2, 2)) / G(1, 3);
                wcg[b + NP * 3 + 3] = (G(3, 3) + G(3, 4) + G(4, 3) + G(4, 4) + 0.5 * (G(2, 3) + G(2, 4) + G(3, 2) + G(4, 2)) + 0.25 * G(2, 2)) / G(3, 3);
                auto fac = [&](double Q, double &f_out, double &sum_out) {
                    std::vector<double> y(n_sls);
                    for (int k = 0; k < n_sls; k++) y[k] = A.y_j[k] / Q;
                    const std::vector<double> yp = A.do_corr_lowq ? fast_correct(y) : y;
                    double s1 = 0.0, s2 = 0.0;
                    for (int k = 0; k < n_sls; k++) { s1 += yp[k] * A.w_j[k] * A.w_j[k] / (w_1 * w_1 + A.w_j[k] * A.w_j[k]); s2 += yp[k]; }
                    f_out = s1 / s2; sum_out = s2;
                };
                double mu_fac, sum_mu, ka_fac, sum_ka;
                fac(ms[r].qmu[e], mu_fac, sum_mu);
                fac(ms[r].qka[e], ka_fac, sum_ka);
                for (int q = 0; q < NPT; q++) {
                    const size_t p = b + q;
                    const double mu = ms[r].mu[p], lam = ms[r].lam[p];
                    const double mu_w1 = mu * (1.0 + 2.0 / (PI * ms[r].qmu[e]) * std::log(w_1 / w_0));
                    const double ka_w1 = (lam + 2.0 / 3.0 * mu) * (1.0 + 2.0 / (PI * ms[r].qka[e]) * std::log(w_1 / w_0));
                    const double dmu0 = mu_w1 / (1.0 / sum_mu + 1.0 - mu_fac), dka0 = ka_w1 / (1.0 / sum_ka + 1.0 - ka_fac);
                    const double w = A.coarse_grained ? wcg[p] : 1.0;
                    mu_u[p] = mu_w1 + w * dmu0 * mu_fac;
                    lam_u[p] = ka_w1 + w * dka0 * ka_fac - 2.0 / 3.0 * mu_u[p];
                    dmu[p] = w * dmu0;
                    dka[p] = w * dka0;
                }
            }
            m.put("attenuation%att_coarse_grained", scalar_i(A.coarse_grained));
            m.put("attenuation%n_sls_attenuation", scalar_i(n_sls));
            m.put("attenuation%do_corr_lowq", scalar_i(A.do_corr_lowq));
            m.put("attenuation%y_j", make(Array::F64, {(uint64_t)n_sls}, A.y_j.data()));
            m.put("attenuation%exp_w_j_deltat", make(Array::F64, {(uint64_t)n_sls}, exp_w.data()));
            m.put("attenuation%ts_fac_t", make(Array::F64, {(uint64_t)n_sls}, ts_t.data()));
            m.put("attenuation%ts_fac_tm1", make(Array::F64, {(uint64_t)n_sls}, ts_tm1.data()));
            m.put("data_matr%Q_mu", f32_of(ms[r].qmu, {(uint64_t)nel_s}));
            m.put("data_matr%Q_kappa", f32_of(ms[r].qka, {(uint64_t)nel_s}));
            if (A.coarse_grained) {
                m.put("data_matr%delta_mu_cg4", f32_of(cg4(dmu, nel_s), {(uint64_t)nel_s, 4}));
                m.put("data_matr%delta_kappa_cg4", f32_of(cg4(dka, nel_s), {(uint64_t)nel_s, 4}));
                for (const char *k : {"DsDeta", "DzDeta", "DsDxi", "DzDxi"}) {
                    // cg4 samples of the (already real(4)) pointwise planes
                    const float *pl = m.f(std::string("data_pointwise%") + k + "_over_J_sol");
                    std::vector<double> tmp((size_t)NPT * nel_s);
                    for (size_t p = 0; p < tmp.size(); p++) tmp[p] = pl[p];
                    m.put(std::string("attenuation%") + k + "_over_J_sol_cg4", f32_of(cg4(tmp, nel_s), {(uint64_t)nel_s, 4}));
                }
            } else {
                put25("data_matr%delta_mu", dmu, nel_s);
                put25("data_matr%delta_kappa", dka, nel_s);
            }
        }
        // ---- stiffness planes ----------------------------------------------------------------------
        {
            Planes P, P0, Pcg;
            solid_stiffness_terms(src, g, ms[r], lam_u, mu_u, opt.attenuation, sp, P, P0, Pcg);
            for (const auto &kv : P) {
                const bool anel_plane = kv.first == "Y" || kv.first.rfind("V_", 0) == 0;
                if (anel_plane && opt.att.coarse_grained) continue;           // only their cg4 samples are kept
                put25("data_matr%" + kv.first, kv.second, nel_s);
            }
            for (const auto &kv : P0) {
                const bool anel_vec = kv.first == "Y0" || kv.first.rfind("V0_", 0) == 0;
                if (anel_vec && opt.att.coarse_grained) continue;
                m.put("data_matr%" + kv.first, f32_of(kv.second, {(uint64_t)nel_s, NP}));
            }
            if (opt.attenuation && opt.att.coarse_grained)
                for (const auto &kv : Pcg) m.put("data_matr%" + kv.first, f32_of(kv.second, {(uint64_t)nel_s, 4}));
        }
        if (nel_f) {
            // def_precomp_terms.f90:2336-2470
            const size_t n = f.s.size();
            std::vector<double> m1(n), m2(n), m4(n), mw(n, 0.0), m0((size_t)NP * nel_f, 0.0), inv_rho(n), fsm(n, 1.0);
            for (int e = 0; e < nel_f; e++)
                for (int q = 0; q < NPT; q++) {
                    const size_t p = (size_t)NPT * e + q;
                    const int i = q % NP, j = q / NP;
                    const double ij = 1.0 / f.jac[p], W = f.W[p], rho = mf[r].rho[p];
                    const double alpha = -ij * f.dsdxi[p] * f.dsdeta[p] * W, beta = ij * f.dsdxi[p] * f.dsdxi[p] * W;
                    const double gamma = ij * f.dsdeta[p] * f.dsdeta[p] * W, delta = -ij * f.dzdxi[p] * f.dzdeta[p] * W;
                    const double epsil = ij * f.dzdxi[p] * f.dzdxi[p] * W, zeta = ij * f.dzdeta[p] * f.dzdeta[p] * W;
                    m1[p] = (delta + alpha) / rho; m2[p] = (zeta + gamma) / rho; m4[p] = (epsil + beta) / rho;
                    inv_rho[p] = 1.0 / rho;
                    if (f.r[p] > router - 1.0) fsm[p] = 0.0;                    // time_evol_wave.F90:1615-1630
                    if (src != "monopole") {
                        const double k4 = src == "quadpole" ? 4.0 : 1.0;
                        const bool i0 = f.axis[e] && i == 0;
                        mw[p] = i0 ? 0.0 : k4 * (f.kwts2[p] / rho);
                        if (i0) m0[(size_t)NP * e + j] = k4 * (f.kwts2[p] / rho);
                    }
                }
            put25("data_matr%M1chi_fl", m1, nel_f);
            put25("data_matr%M2chi_fl", m2, nel_f);
            put25("data_matr%M4chi_fl", m4, nel_f);
            if (src != "monopole") {
                put25("data_matr%M_w_fl", mw, nel_f);
                m.put("data_matr%M0_w_fl", f32_of(m0, {(uint64_t)nel_f, NP}));
            }
            put25("data_matr%inv_rho_fluid", inv_rho, nel_f);
            put25("data_mesh%fluid_free_surface_mask", fsm, nel_f);
        }
        // ---- S/F boundary terms (def_precomp_terms.f90:2474-2712) -----------------------------------
        const int nel_bdry = m.int_of("data_mesh%nel_bdry");
        if (nel_bdry > 0 && m.int_of("data_mesh%have_bdry_elem", 1)) {
            const int32_t *bs = m.i("data_mesh%bdry_solid_el"), *bjs = m.i("data_mesh%bdry_jpol_solid");
            const int32_t *bf = m.i("data_mesh%bdry_fluid_el");
            std::vector<double> bm((size_t)2 * nel_bdry * NP, 0.0);
            double bdry_sum = 0.0;
            std::vector<double> bdry_radius(nel_bdry);
            for (int b = 0; b < nel_bdry; b++) {
                const int e = bs[b] - 1, jj = bjs[b];
                const size_t p0 = (size_t)NPT * e + NP * jj;
                auto theta_of = [&](size_t p) { return std::atan2(g.s[p], g.z[p]); };
                const double th1 = theta_of(p0), th2 = theta_of(p0 + 4), delta_th = 0.5 * std::fabs(th2 - th1);
                const double rr = g.r[p0 + 4];
                double *b1 = &bm[(size_t)b * NP], *b2 = &bm[((size_t)nel_bdry + b) * NP];
                if (g.axis[e]) {
                    for (int i = 1; i < NP; i++) {
                        const double th = theta_of(p0 + i);
                        b1[i] = delta_th * sp.wt_axial_k[i] * std::sin(th) / (1.0 + sp.xi_k[i]) * std::sin(th);
                        b2[i] = delta_th * sp.wt_axial_k[i] * std::sin(th) / (1.0 + sp.xi_k[i]) * std::cos(th);
                        bdry_sum += delta_th * sp.wt_axial_k[i] * std::sin(th) / (1.0 + sp.xi_k[i]);
                    }
                    bdry_sum += 1.0 / rr * delta_th * sp.wt_axial_k[0] * g.dsdxi[p0];
                    b1[0] = 0.0;
                    // (the reference has cos(0) = 1 here also at the southern axis, :2603)
                    b2[0] = 1.0 / rr * delta_th * sp.wt_axial_k[0] * g.dsdxi[p0];
                } else {
                    for (int i = 0; i < NP; i++) {
                        const double th = theta_of(p0 + i);
                        b1[i] = delta_th * sp.wt[i] * std::sin(th) * std::sin(th);
                        b2[i] = delta_th * sp.wt[i] * std::sin(th) * std::cos(th);
                        bdry_sum += delta_th * sp.wt[i] * std::sin(th);
                    }
                }
                // fluid above the solid (e.g. the ICB): negative (:2686-2696)
                const double r_fl = f.r_mid[bf[b] - 1], r_so = g.r_mid[e];
                const double sign = r_so > r_fl ? 1.0 : -1.0;
                for (int i = 0; i < NP; i++) { b1[i] *= sign * rr * rr; b2[i] *= sign * rr * rr; }
                bdry_radius[b] = rr;
            }
            m.put("data_matr%bdry_matr", f32_of(bm, {2, (uint64_t)nel_bdry, NP}));
            m.put("precomp%bdry_sum", scalar_d(bdry_sum));
            m.put("precomp%solflubdry_radius", make(Array::F64, {(uint64_t)nel_bdry}, bdry_radius.data()));
        }
        // ---- source (source.f90:454-476, 921-1226): point source on the northern axis ----------------
        {
            std::vector<float> st((size_t)3 * 8 * NPT, 0.0f);
            std::vector<int32_t> ielsrc(8, 0);
            int nelsrc = 0;
            const double zsrc = router - opt.src_depth;
            // find_srcloc: the on-axis GLL point closest in z, ties keep a second element
            double dmin = 1e300;
            for (int e = 0; e < nel_s; e++)
                if (g.axis[e])
                    for (int j = 0; j < NP; j++) {
                        const size_t p = (size_t)NPT * e + NP * j;
                        if (g.z[p] > 0) dmin = std::min(dmin, std::fabs(g.z[p] - zsrc));
                    }
            // is the source in this rank?  (only the rank that holds the closest point over all ranks)
            double dmin_all = dmin;
            for (size_t q = 0; q < nr; q++) {
                const Geometry &G = gs[q];
                for (int e = 0; e < G.nel; e++)
                    if (G.axis[e])
                        for (int j = 0; j < NP; j++) {
                            const size_t p = (size_t)NPT * e + NP * j;
                            if (G.z[p] > 0) dmin_all = std::min(dmin_all, std::fabs(G.z[p] - zsrc));
                        }
            }
            std::vector<std::pair<int, int>> srcs;       // (element, jpol)
            if (dmin <= dmin_all * (1 + 1e-12) + 1e-6)
                for (int e = 0; e < nel_s && srcs.size() < 2; e++)
                    if (g.axis[e])
                        for (int j = 0; j < NP; j++) {
                            const size_t p = (size_t)NPT * e + NP * j;
                            if (g.z[p] > 0 && std::fabs(g.z[p] - zsrc) <= dmin * (1 + 1e-12) + 1e-6 && srcs.size() < 2)
                                srcs.push_back({e, j});
                        }
            if (!srcs.empty()) {
                const int32_t *ig = m.i("data_mesh%igloc_solid");
                std::set<int32_t> gids;
                for (auto &sj : srcs) for (int q = 0; q < NPT; q++) gids.insert(ig[(size_t)NPT * sj.first + q]);
                std::vector<int> cand;
                for (int e = 0; e < nel_s; e++) {
                    bool hit = false;
                    for (int q = 0; q < NPT && !hit; q++) hit = gids.count(ig[(size_t)NPT * e + q]) != 0;
                    if (hit) cand.push_back(e);
                }
                std::map<int, int> loc;
                for (size_t k = 0; k < cand.size(); k++) loc[cand[k]] = (int)k;
                std::vector<double> term(cand.size() * 3 * NPT, 0.0);      // [cand][comp][j][i]
                const bool force = opt.src_type2 == "vertforce" || opt.src_type2 == "thetaforce" || opt.src_type2 == "phiforce";
                const float *G1T = m.f("data_spec%G1T"), *G2T = m.f("data_spec%G2T"), *G2 = m.f("data_spec%G2");
                const float *pdze = m.f("data_pointwise%DzDeta_over_J_sol"), *pdzx = m.f("data_pointwise%DzDxi_over_J_sol");
                const float *pdse = m.f("data_pointwise%DsDeta_over_J_sol"), *pdsx = m.f("data_pointwise%DsDxi_over_J_sol");
                for (auto &sj : srcs) {
                    const int e = sj.first, jp = sj.second, ip = 0, q0 = loc[e];
                    double *T = &term[(size_t)q0 * 3 * NPT];
                    if (force) { T[(opt.src_type2 == "vertforce" ? 2 : 0) * NPT + NP * jp + ip] = 1.0; continue; }
                    const float *GT = g.axis[e] ? G1T : G2T;               // Fortran G(a,b) at [a + 5 b]
                    const size_t ps = (size_t)NPT * e + NP * jp + ip;
                    for (int ipol = 0; ipol < NP; ipol++)
                        for (int jpol = 0; jpol < NP; jpol++) {
                            // ws = unit field at (ipol, jpol): mxm1(ip,jp) = GT(ip,ipol) [jpol == jp], mxm2 = G2(jpol,jp) [ipol == ip]
                            const double mxm1 = jpol == jp ? (double)GT[ip + NP * ipol] : 0.0;
                            const double mxm2 = ipol == ip ? (double)G2[jpol + NP * jp] : 0.0;
                            const double dsws = (double)pdze[ps] * mxm1 + (double)pdzx[ps] * mxm2;
                            const double dzwz = (double)pdse[ps] * mxm1 + (double)pdsx[ps] * mxm2;
                            const size_t k = (size_t)NP * jpol + ipol;
                            if (src == "monopole") {
                                if (opt.src_type2 == "explosion") { T[k] = 2.0 * dsws; T[2 * NPT + k] = dzwz; }
                                else if (opt.src_type2 == "mtt_p_mpp") T[k] = dsws;
                                else T[2 * NPT + k] = dzwz;                      // mrr
                            } else if (src == "dipole") { T[k] = dzwz; T[2 * NPT + k] = dsws; }
                            else { T[k] = dsws; T[NPT + k] = dsws; }
                        }
                }
                if (!force) for (double &x : term) x /= (double)srcs.size();
                // assembly over the local mesh (pdistsum_solid, source.f90:1133)
                for (int c = 0; c < 3; c++) {
                    std::map<int32_t, double> sum;
                    for (size_t k = 0; k < cand.size(); k++)
                        for (int q = 0; q < NPT; q++) sum[ig[(size_t)NPT * cand[k] + q]] += term[(k * 3 + c) * NPT + q];
                    for (size_t k = 0; k < cand.size(); k++)
                        for (int q = 0; q < NPT; q++) term[(k * 3 + c) * NPT + q] = sum[ig[(size_t)NPT * cand[k] + q]];
                }
                for (double &x : term) {
                    if (!force && std::fabs(x) < 1e-30) x = 0.0;
                    x /= (src == "dipole" ? PI : 2.0 * PI);
                }
                for (size_t k = 0; k < cand.size(); k++) {
                    if (!force && g.axis[cand[k]])
                        for (int j = 0; j < NP; j++) {
                            double *T = &term[k * 3 * NPT];
                            if (src == "monopole") T[NP * j] = 0.0;
                            else if (src == "dipole") { T[NPT + NP * j] = 0.0; T[2 * NPT + NP * j] = 0.0; }
                            else { T[NP * j] = 0.0; T[NPT + NP * j] = 0.0; T[2 * NPT + NP * j] = 0.0; }
                        }
                    double mx = 0.0;
                    for (int q = 0; q < 3 * NPT; q++) mx = std::max(mx, std::fabs(term[k * 3 * NPT + q]));
                    if (mx > 0) {
                        if (nelsrc >= 8) throw SolverError("more than 8 source elements");
                        ielsrc[nelsrc] = cand[k] + 1;
                        for (int c = 0; c < 3; c++)
                            for (int q = 0; q < NPT; q++) st[((size_t)c * 8 + nelsrc) * NPT + q] = (float)term[(k * 3 + c) * NPT + q];
                        nelsrc++;
                    }
                }
            }
            m.put("data_source%have_src_in_fluid", scalar_i(0));
            m.put("data_source%nelsrc", scalar_i(nelsrc));
            m.put("data_source%ielsrc", i32_of(ielsrc, {8}));
            m.put("data_source%source_term_el", make(Array::F32, {3, 8, NP, NP}, st.data()));
            std::vector<float> stf(std::max(opt.niter, 1), 0.0f);
            if (opt.time_scheme == "newmark2" && opt.niter > 0) stf = compute_stf(opt, opt.niter, deltat);
            else if (opt.stf_type == "dirac_1") throw SolverError("source time function non existant for the symplectic schemes: dirac_1");
            m.put("data_source%stf", make(Array::F32, {(uint64_t)stf.size()}, stf.data()));
            m.put("data_source%stf_type", scalar_i(stf_code(opt.stf_type)));
            m.put("data_source%decay", scalar_d(opt.decay));
            m.put("data_source%t_0", scalar_d(opt.t_0));
            m.put("data_source%shift_fact", scalar_d(stf_shift(opt, deltat)));
            m.put("data_source%magnitude", scalar_d(opt.magnitude));
            {
                std::vector<int32_t> chars(opt.src_type2.begin(), opt.src_type2.end()), stfc(opt.stf_type.begin(), opt.stf_type.end());
                m.put("data_source%src_type2", i32_of(chars, {(uint64_t)chars.size()}));
                m.put("data_source%stf_name", i32_of(stfc, {(uint64_t)stfc.size()}));
                m.put("data_source%src_depth", scalar_d(opt.src_depth));
            }
        }
        // ---- receivers: nearest surface GLL point to each colatitude, owned by the rank that holds it ---
        {
            std::vector<int32_t> rec;       // (num_rec, 3) Fortran order, filled later
            std::vector<std::array<int32_t, 3>> hits;
            std::vector<int32_t> loc2glob;   // loc2globrec (1-based position in the receiver list)
            std::vector<double> th_deg;      // recfile_th: colatitude of the grid point taken [deg]
            for (size_t irec = 0; irec < opt.rec_colat_deg.size(); irec++) {
                const double cd = opt.rec_colat_deg[irec];
                const double c = cd * PI / 180.0;
                double th_best = 0.0;
                // every rank finds its closest surface point (first one wins within the rank); of the ranks that
                // reach the global minimum the one with the largest number takes the receiver (seismograms.f90:477-484)
                double best = 1e300;
                size_t best_rank = nr;
                std::array<int32_t, 3> at{0, 0, 0};
                for (size_t q = 0; q < nr; q++) {
                    const Geometry &G = gs[q];
                    double mine = 1e300, th_mine = 0.0;
                    std::array<int32_t, 3> at_mine{0, 0, 0};
                    for (int e = 0; e < G.nel; e++)
                        for (int k = 0; k < NPT; k++) {
                            const size_t p = (size_t)NPT * e + k;
                            if (std::fabs(G.r[p] - router) > 1e-3 * router * 1e-3) continue;
                            const double d = std::fabs(std::atan2(G.s[p], G.z[p]) - c);
                            if (d < mine - 1e-14) { mine = d; at_mine = {e + 1, k % NP, k / NP}; th_mine = std::atan2(G.s[p], G.z[p]); }
                        }
                    if (mine < 1e299 && (mine < best - 1e-12 || std::fabs(mine - best) <= 1e-12)) {
                        best = std::fmin(best, mine); best_rank = q; at = at_mine; th_best = th_mine;
                    }
                }
                if (best_rank == r) { hits.push_back(at); loc2glob.push_back((int32_t)irec + 1); th_deg.push_back(th_best * 180.0 / PI); }
            }
            const size_t nrec = hits.size();
            rec.resize(3 * nrec);
            for (size_t k = 0; k < nrec; k++) { rec[k] = hits[k][0]; rec[nrec + k] = hits[k][1]; rec[2 * nrec + k] = hits[k][2]; }
            m.put("data_mesh%num_rec", scalar_i((int32_t)nrec));
            m.put("data_mesh%recfile_el", i32_of(rec, {3, (uint64_t)nrec}));
            m.put("data_mesh%loc2globrec", i32_of(loc2glob, {(uint64_t)nrec}));
            m.put("data_mesh%recfile_th", make(Array::F64, {(uint64_t)nrec}, th_deg.data()));
            m.put("data_mesh%recfile_readth", make(Array::F64, {(uint64_t)opt.rec_colat_deg.size()}, opt.rec_colat_deg.data()));   // all receivers of the run
        }
        // ---- wavefield-dump point set (meshes_io.F90:489-640: first visit wins, solid then fluid) -------
        m.put("data_io%dump_wavefields", scalar_i(opt.dump_wavefields && opt.strain_it > 0));
        // ---- xdmf plot points (dump_xdmf_grid, meshes_io.F90:110-437): elements with a corner inside the
        // plot region, their (i_arr, j_arr) points de-duplicated by global number, fluid elements first ----
        m.put("data_io%dump_xdmf", scalar_i(opt.snap_it > 0));
        if (opt.snap_it > 0) {
            const int in = (int)opt.xdmf_gll_i.size(), jn = (int)opt.xdmf_gll_j.size();
            const int nelem = nel_f + nel_s;
            std::vector<int32_t> mask((size_t)in * jn * nelem, 0), map((size_t)in * jn * nelem, 0);
            std::vector<char> in_range(nelem, 0);
            int nin = 0;
            for (int d = 0; d < 2; d++) {
                const Geometry &G = d ? g : f;                 // d = 0: fluid
                const int nel = d ? nel_s : nel_f, off = d ? nel_f : 0;
                for (int e = 0; e < nel; e++) {
                    double rmin = 1e300, rmax = -1e300, tmin = 1e300, tmax = -1e300;
                    for (int q : {0, NP - 1, NPT - NP, NPT - 1}) {
                        const size_t p = (size_t)NPT * e + q;
                        const double th = std::atan2(G.s[p], G.z[p]);
                        rmin = std::min(rmin, G.r[p]); rmax = std::max(rmax, G.r[p]);
                        tmin = std::min(tmin, th); tmax = std::max(tmax, th);
                    }
                    if (rmin < opt.xdmf_rmax && rmax > opt.xdmf_rmin && tmin < opt.xdmf_thetamax && tmax > opt.xdmf_thetamin) {
                        in_range[off + e] = 1;
                        nin++;
                    }
                }
            }
            const int nelem_plot = nin * (in - 1) * (jn - 1);
            std::map<int64_t, int32_t> seen;
            std::vector<double> points;
            for (int d = 0; d < 2; d++) {
                const Geometry &G = d ? g : f;
                const int nel = d ? nel_s : nel_f, off = d ? nel_f : 0;
                const int32_t *ig = m.i(d ? "data_mesh%igloc_solid" : "data_mesh%igloc_fluid");
                const int64_t gofs = d ? m.int_of("data_mesh%nglob_fluid") : 0;
                for (int e = 0; e < nel; e++) {
                    if (!in_range[off + e]) continue;
                    for (int i = 0; i < in; i++)
                        for (int j = 0; j < jn; j++) {
                            const size_t p = (size_t)NPT * e + opt.xdmf_gll_i[i] + NP * opt.xdmf_gll_j[j];
                            const size_t k = i + (size_t)in * (j + (size_t)jn * (off + e));
                            auto it = seen.find(ig[p] + gofs);
                            if (it == seen.end()) {
                                it = seen.emplace(ig[p] + gofs, (int32_t)seen.size() + 1).first;
                                mask[k] = 1;
                                points.push_back(G.s[p]);
                                points.push_back(G.z[p]);
                            }
                            map[k] = it->second;
                        }
                }
            }
            std::vector<int32_t> grid;
            grid.reserve((size_t)4 * nelem_plot);
            for (int el = 0; el < nelem; el++) {
                if (!in_range[el]) continue;
                auto at = [&](int i, int j) { return map[i + (size_t)in * (j + (size_t)jn * el)] - 1; };
                for (int i = 0; i < in - 1; i++)
                    for (int j = 0; j < jn - 1; j++)
                        for (int32_t c : {at(i, j), at(i + 1, j), at(i + 1, j + 1), at(i, j + 1)}) grid.push_back(c);
            }
            m.put("data_time%snap_it", scalar_i(opt.snap_it));
            m.put("data_io%i_arr_xdmf", i32_of(std::vector<int32_t>(opt.xdmf_gll_i.begin(), opt.xdmf_gll_i.end()), {(uint64_t)in}));
            m.put("data_io%j_arr_xdmf", i32_of(std::vector<int32_t>(opt.xdmf_gll_j.begin(), opt.xdmf_gll_j.end()), {(uint64_t)jn}));
            m.put("data_mesh%plotting_mask", i32_of(mask, {(uint64_t)nelem, (uint64_t)jn, (uint64_t)in}));
            m.put("data_mesh%mapping_ijel_iplot", i32_of(map, {(uint64_t)nelem, (uint64_t)jn, (uint64_t)in}));
            m.put("data_mesh%npoint_plot", scalar_i((int32_t)seen.size()));
            m.put("data_mesh%nelem_plot", scalar_i(nelem_plot));
            m.put("data_mesh%xdmf_points", f32_of(points, {(uint64_t)seen.size(), 2}));
            m.put("data_mesh%xdmf_grid", i32_of(grid, {(uint64_t)nelem_plot, 4}));
        }
        if (opt.dump_wavefields && opt.strain_it > 0) {
            const size_t n = (size_t)NPT * (nel_s + nel_f);
            std::vector<int32_t> mask(n, 0), map(n, 0);
            int base = 0, counts[2] = {0, 0};
            for (int d = 0; d < 2; d++) {
                const int nel = d ? nel_f : nel_s;
                const int32_t *ig = m.i(d ? "data_mesh%igloc_fluid" : "data_mesh%igloc_solid");
                std::map<int32_t, int32_t> first;
                const size_t off = d ? (size_t)NPT * nel_s : 0;
                for (size_t p = 0; p < (size_t)NPT * nel; p++) {
                    auto it = first.find(ig[p]);
                    if (it == first.end()) { it = first.emplace(ig[p], (int32_t)first.size() + 1).first; mask[off + p] = 1; }
                    map[off + p] = it->second + base;
                }
                counts[d] = (int)first.size();
                base += counts[d];
            }
            m.put("data_mesh%kwf_mask", i32_of(mask, {(uint64_t)(nel_s + nel_f), NP, NP}));
            m.put("data_mesh%mapping_ijel_ikwf", i32_of(map, {(uint64_t)(nel_s + nel_f), NP, NP}));
            m.put("data_mesh%npoint_solid_kwf", scalar_i(counts[0]));
            m.put("data_mesh%npoint_fluid_kwf", scalar_i(counts[1]));
            m.put("data_io%dump_type", scalar_i(0));
            // ---- the Mesh group of the output database (nc_routines.F90:668-823, meshes_io.F90:641-778):
            // coordinates and the model at the dumped points, and per dumped element its mid-point, corner and
            // GLL point numbers in that list (0-based within the rank)
            const int npt = counts[0] + counts[1];
            std::vector<double> S(npt, 0.0), Z(npt, 0.0);
            std::vector<float> vp(npt, 0.f), vs(npt, 0.f), rho(npt, 0.f), lam(npt, 0.f), mu(npt, 0.f), xi(npt, 0.f), phi(npt, 0.f),
                eta(npt, 0.f), qmu(npt, 0.f), qka(npt, 0.f);
            const int nel = nel_s + nel_f;
            std::vector<int32_t> mid(nel), elt(nel), axs(nel), fem((size_t)4 * nel), sem((size_t)NPT * nel);
            std::vector<double> mps(nel), mpz(nel);
            const int32_t *eltype = m.i("data_mesh%eltype");
            for (int d = 0; d < 2; d++) {
                const Geometry &G = d ? f : g;
                const Material &M = d ? mf[r] : ms[r];
                const int ne = d ? nel_f : nel_s;
                const size_t off = d ? (size_t)NPT * nel_s : 0;
                for (int e = 0; e < ne; e++) {
                    for (int q = 0; q < NPT; q++) {
                        const size_t p = (size_t)NPT * e + q;
                        const int k = map[off + p] - 1;
                        sem[off + p] = k;
                        if (!mask[off + p]) continue;
                        S[k] = G.s[p]; Z[k] = G.z[p];
                        rho[k] = (float)M.rho[p]; lam[k] = (float)M.lam[p]; mu[k] = (float)M.mu[p];
                        xi[k] = (float)M.xi[p]; phi[k] = (float)M.phi[p]; eta[k] = (float)M.eta[p];
                        vp[k] = (float)std::sqrt((M.lam[p] + 2.0 * M.mu[p]) / M.rho[p]);
                        vs[k] = (float)std::sqrt(M.mu[p] / M.rho[p]);
                        qmu[k] = (float)M.qmu[e]; qka[k] = (float)M.qka[e];
                    }
                    const size_t ee = (d ? nel_s : 0) + e, p0 = off + (size_t)NPT * e;
                    mid[ee] = map[p0 + 2 * NP + 2] - 1;
                    elt[ee] = eltype[G.iel_glob[e]];
                    axs[ee] = G.axis[e];
                    fem[4 * ee] = map[p0] - 1; fem[4 * ee + 1] = map[p0 + NP - 1] - 1;
                    fem[4 * ee + 2] = map[p0 + NPT - 1] - 1; fem[4 * ee + 3] = map[p0 + NPT - NP] - 1;
                    mps[ee] = G.s[(size_t)NPT * e + 2 * NP + 2]; mpz[ee] = G.z[(size_t)NPT * e + 2 * NP + 2];
                }
            }
            const uint64_t un = (uint64_t)npt, ue = (uint64_t)nel;
            m.put("nc_mesh%mesh_S", make(Array::F64, {un}, S.data()));
            m.put("nc_mesh%mesh_Z", make(Array::F64, {un}, Z.data()));
            const std::pair<const char *, const std::vector<float> *> fl[] = {{"vp", &vp}, {"vs", &vs}, {"rho", &rho}, {"lambda", &lam},
                {"mu", &mu}, {"xi", &xi}, {"phi", &phi}, {"eta", &eta}, {"Qmu", &qmu}, {"Qka", &qka}};
            for (const auto &kv : fl) m.put(std::string("nc_mesh%mesh_") + kv.first, make(Array::F32, {un}, kv.second->data()));
            m.put("nc_mesh%midpoint_mesh", i32_of(mid, {ue}));
            m.put("nc_mesh%eltype", i32_of(elt, {ue}));
            m.put("nc_mesh%axis", i32_of(axs, {ue}));
            m.put("nc_mesh%fem_mesh", i32_of(fem, {ue, 4}));
            m.put("nc_mesh%sem_mesh", i32_of(sem, {ue, NP, NP}));
            m.put("nc_mesh%mp_mesh_S", make(Array::F64, {ue}, mps.data()));
            m.put("nc_mesh%mp_mesh_Z", make(Array::F64, {ue}, mpz.data()));
        }
        // ---- time ------------------------------------------------------------------------------------------
        static const std::map<std::string, int> SCHEMES = {{"newmark2", 0}, {"symplec4", 1}, {"ML_SO4m5", 2},
                                                           {"ML_SO6m7", 3}, {"KL_O8m17", 4}, {"SS_35o10", 5}};
        const auto sc = SCHEMES.find(opt.time_scheme);
        if (sc == SCHEMES.end()) throw SolverError("unknown time scheme " + opt.time_scheme);
        m.put("data_time%time_scheme", scalar_i(sc->second));
        m.put("data_time%deltat", scalar_d(deltat));
        m.put("data_time%niter", scalar_i(opt.niter));
        m.put("data_time%seis_it", scalar_i(opt.seis_it));
        m.put("data_time%strain_it", scalar_i(opt.strain_it));
        // what the checks below need
        {
            double vs = 0.0, vf = 0.0;
            for (double x : g.massmat_k) vs += x;
            for (double x : f.massmat_k) vf += x;
            m.put("precomp%solid_volume_over_2pi", scalar_d(vs));
            m.put("precomp%fluid_volume_over_2pi", scalar_d(vf));
        }
    }
}

// ---- the SLS fit (invert_linear_solids + q_linear_solid + l2_error, attenuation.f90:1099-1339) ---------------------
std::vector<double> q_linear_solid(const std::vector<double> &y_j, const std::vector<double> &w_j, const std::vector<double> &w) {
    std::vector<double> q(w.size());
    for (size_t k = 0; k < w.size(); k++) {
        double den = 0.0;
        for (size_t j = 0; j < y_j.size(); j++) den += y_j[j] * w[k] * w_j[j] / (w[k] * w[k] + w_j[j] * w_j[j]);
        q[k] = 1.0 / den;
    }
    return q;
}

double fit_linear_solids(AttenuationOptions &A, int n_sls, double f_min, double f_max, uint64_t seed, int max_it) {
    const int nfsamp = 100;
    double Tw = 0.1, Ty = 0.1;
    const double d = 0.99995;
    std::vector<double> w_j(n_sls), y_j(n_sls, 1.5), w(nfsamp), weights(nfsamp);
    if (n_sls > 1) {
        const double expo = (std::log10(f_max) - std::log10(f_min)) / (n_sls - 1.0);
        for (int j = 0; j < n_sls; j++) w_j[j] = 2 * PI * std::pow(10.0, std::log10(f_min) + j * expo);
    } else w_j[0] = std::sqrt(f_max * f_min) * 2 * PI;
    const double expo = (std::log10(f_max) - std::log10(f_min)) / (nfsamp - 1.0);
    double wsum = 0.0;
    for (int k = 0; k < nfsamp; k++) { w[k] = 2 * PI * std::pow(10.0, std::log10(f_min) + k * expo); wsum += w[k]; }
    for (int k = 0; k < nfsamp; k++) weights[k] = w[k] / wsum * nfsamp;          // FREQ_WEIGHT true
    auto misfit = [&](const std::vector<double> &y, const std::vector<double> &wj) {
        const std::vector<double> q = q_linear_solid(y, wj, w);
        double lse = 0.0;
        for (int k = 0; k < nfsamp; k++) { const double l = std::log(1.0 / q[k]); lse += l * l * weights[k]; }    // Q_target = 1
        return std::sqrt(lse / (double)nfsamp);
    };
    double chi = misfit(y_j, w_j);
    // the reference draws from an unseeded random_number; a seeded 64-bit generator makes the fit repeatable
    uint64_t s = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    std::vector<double> wt(n_sls), yt(n_sls);
    for (int it = 0; it < max_it; it++) {
        for (int j = 0; j < n_sls; j++) {
            wt[j] = w_j[j] * (1.0 + (0.5 - rnd()) * Tw);
            yt[j] = y_j[j] * (1.0 + (0.5 - rnd()) * Ty);
        }
        const double c = misfit(yt, wt);
        Tw *= d; Ty *= d;
        if (c < chi) { y_j = yt; w_j = wt; chi = c; }
    }
    A.w_j = w_j; A.y_j = y_j; A.f_min = f_min; A.f_max = f_max;
    return chi;
}

std::vector<std::vector<int>> receiver_indices(const std::vector<Modules> &ranks) {
    std::vector<std::vector<int>> out;
    for (const Modules &m : ranks) {
        const size_t n = (size_t)m.int_of("data_mesh%num_rec");
        const int32_t *p = n ? m.i("data_mesh%loc2globrec") : nullptr;
        out.emplace_back(p, p + n);
    }
    return out;
}
std::vector<std::vector<double>> receiver_colatitudes(const std::vector<Modules> &ranks) {
    std::vector<std::vector<double>> out;
    for (const Modules &m : ranks) {
        const size_t n = (size_t)m.int_of("data_mesh%num_rec");
        const double *p = n ? m.d("data_mesh%recfile_th") : nullptr;
        out.emplace_back(p, p + n);
    }
    return out;
}

PrecompChecks precompute_checks(const std::vector<Modules> &ranks) {
    PrecompChecks c{0, 0, 0, 0, 0, 0};
    std::set<long> radii;
    for (const Modules &m : ranks) {
        c.solid_volume += 2 * PI * m.real_of("precomp%solid_volume_over_2pi");
        c.fluid_volume += 2 * PI * m.real_of("precomp%fluid_volume_over_2pi");
        const double router = m.real_of("data_mesh%router"), rmin = m.real_of("data_mesh%rmin");
        c.sphere_volume = 4.0 / 3.0 * PI * router * router * router;
        c.hollow_volume = 4.0 / 3.0 * PI * rmin * rmin * rmin;
        if (m.has("precomp%bdry_sum")) {
            // int sin(theta) dtheta over [0, pi] = 2 per S/F boundary (def_precomp_terms.f90:2743)
            c.bdry_sum += m.real_of("precomp%bdry_sum");
            const Array &rr = m.at("precomp%solflubdry_radius");
            for (size_t k = 0; k < rr.count(); k++) radii.insert(std::lround(rr.f64()[k]));
        }
    }
    c.n_sf_boundaries = (int)radii.size();
    return c;
}

}  // namespace axisem
