#include "background_models.hpp"

#include <cmath>
#include <stdexcept>

namespace axisem {
namespace {

using V = std::vector<double>;

ModelDomain iso(double r0, double r1, bool fluid, double qmu, double qka, V rho, V vp, V vs) {
    return ModelDomain{r0, r1, fluid, qmu, qka, rho, vp, vs, vp, vs, V{1.0}};
}

// From the surface inwards.  Below 6151 km (220 km depth) PREM is isotropic in both variants.
std::vector<ModelDomain> build(bool anisotropic) {
    std::vector<ModelDomain> d;
    d.push_back(iso(6356.0, 6371.0, false, 600.0, 57827.0, {2.6}, {5.8}, {3.2}));                 // 1 upper crust
    d.push_back(iso(6346.6, 6356.0, false, 600.0, 57827.0, {2.9}, {6.8}, {3.9}));                 // 2 lower crust
    const V rho_um{2.6910, 0.6924};
    const double q_um[2] = {600.0, 80.0};                                                            // 3 LID, 4 LVZ
    const double r_um[3] = {6346.6, 6291.0, 6151.0};
    for (int k = 0; k < 2; k++) {
        if (anisotropic)
            d.push_back(ModelDomain{r_um[k + 1], r_um[k], false, q_um[k], 57827.0, rho_um,
                                    {0.8317, 7.2180}, {5.8582, -1.4678}, {3.5908, 4.6172}, {-1.0839, 5.7176},
                                    {3.3687, -2.4778}});
        else
            d.push_back(iso(r_um[k + 1], r_um[k], false, q_um[k], 57827.0, rho_um, {4.1875, 3.9382}, {2.1519, 2.3481}));
    }
    d.push_back(iso(5971.0, 6151.0, false, 143.0, 57827.0, {7.1089, -3.8045}, {20.3926, -12.2569}, {8.9496, -4.4597}));     // 5
    d.push_back(iso(5771.0, 5971.0, false, 143.0, 57827.0, {11.2494, -8.0298}, {39.7027, -32.6166}, {22.3512, -18.5856}));  // 6
    d.push_back(iso(5701.0, 5771.0, false, 143.0, 57827.0, {5.3197, -1.4836}, {19.0957, -9.8672}, {9.9839, -4.9324}));      // 7
    const V rho_lm{7.9565, -6.4761, 5.5283, -3.0807};
    d.push_back(iso(5600.0, 5701.0, false, 312.0, 57827.0, rho_lm, {29.2766, -23.6027, 5.5242, -2.5514},
                    {22.3459, -17.2473, -2.0834, 0.9783}));                                                                  // 8
    d.push_back(iso(3630.0, 5600.0, false, 312.0, 57827.0, rho_lm, {24.9520, -40.4673, 51.4832, -26.6419},
                    {11.1671, -13.7818, 17.4575, -9.2777}));                                                                 // 9
    d.push_back(iso(3480.0, 3630.0, false, 312.0, 57827.0, rho_lm, {15.3891, -5.3181, 5.5242, -2.5514},
                    {6.9254, 1.4672, -2.0834, 0.9783}));                                                                     // 10
    d.push_back(iso(1221.5, 3480.0, true, 0.0, 57827.0, {12.5815, -1.2638, -3.6426, -5.5281},
                    {11.0487, -4.0362, 4.8023, -13.5732}, {0.0}));                                                           // 11 outer core
    d.push_back(iso(0.0, 1221.5, false, 84.6, 1327.7, {13.0885, 0.0, -8.8381}, {11.2622, 0.0, -6.3640},
                    {3.6678, 0.0, -4.4475}));                                                                                // 12 inner core
    return d;
}

double poly(const V &c, double x) {
    double s = 0.0, xk = 1.0;
    for (double ck : c) { s += ck * xk; xk *= x; }
    return s;
}

}  // namespace

const std::vector<ModelDomain> &model_domains(const std::string &name) {
    static const std::vector<ModelDomain> iso_d = build(false), ani_d = build(true);
    if (name == "prem_iso") return iso_d;
    if (name == "prem_ani") return ani_d;
    throw std::invalid_argument("unknown background model '" + name + "' (prem_iso, prem_ani)");
}

int model_domain_of(const std::string &name, double r_m, bool upper_side) {
    const auto &d = model_domains(name);
    const double r = r_m / 1000.0;
    for (size_t k = 0; k < d.size(); k++) {
        const bool above_bottom = upper_side ? r >= d[k].r_bot_km - 1e-9 : r > d[k].r_bot_km + 1e-9;
        if (above_bottom && r <= d[k].r_top_km + 1e-9) return (int)k + 1;
    }
    if (r <= 1e-9) return (int)d.size();
    throw std::invalid_argument("radius outside the model");
}

ModelValues model_evaluate(const std::string &name, double r_m, int idom) {
    const auto &d = model_domains(name);
    if (idom < 1 || idom > (int)d.size()) throw std::invalid_argument("idom out of range");
    const ModelDomain &m = d[idom - 1];
    const double x = (r_m / 1000.0) / 6371.0;
    return ModelValues{poly(m.rho, x) * 1000.0, poly(m.vpv, x) * 1000.0, poly(m.vsv, x) * 1000.0,
                       poly(m.vph, x) * 1000.0, poly(m.vsh, x) * 1000.0, poly(m.eta, x), m.qmu, m.qkappa};
}

}  // namespace axisem
