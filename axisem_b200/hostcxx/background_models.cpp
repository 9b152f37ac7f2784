#include "background_models.hpp"

#include <cctype>
#include <cmath>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace axisem {
namespace {

using V = std::vector<double>;
const double QKA_INF = 57827.0;

ModelDomain iso(double r0, double r1, bool fluid, V qmu, V qka, V rho, V vp, V vs) {
    return ModelDomain{r0, r1, fluid, qmu, qka, rho, vp, vs, vp, vs, V{1.0}};
}

// ---- PREM and its variants (prem_sub :417, prem_ani_sub :534, prem_solid_sub :1191, prem_onecrust_sub
// :1276, prem_onecrust_ani_sub :1387, prem_light_sub :1524, prem_light_ani_sub :1629,
// prem_solid_light_sub :1759, prem_crust20_{ocean,cont,global}_sub :680 / :850 / :1020) -----------------
struct PremOptions {
    bool anisotropic = false;      // LID and LVZ transversely isotropic
    int crust = 2;                 // 2: upper + lower crust, 1: one crustal layer, 0: mantle up to the surface
    bool solid = false;            // LID and LVZ merged (Q is not defined), outer core with v_s = v_p / sqrt(3)
    const double (*crust20)[4] = nullptr;   // five crustal layers {r_bot, rho, vp, vs} on top of the LID
};

std::vector<ModelDomain> prem(const PremOptions &o) {
    std::vector<ModelDomain> d;
    const V rho_um{2.6910, 0.6924};
    double top = 6371.0;
    if (o.crust20) {
        for (int k = 0; k < 5; k++) {
            d.push_back(iso(o.crust20[k][0], top, false, {600.0}, {QKA_INF}, {o.crust20[k][1]}, {o.crust20[k][2]}, {o.crust20[k][3]}));
            top = o.crust20[k][0];
        }
    } else if (o.crust == 2) {
        d.push_back(iso(6356.0, 6371.0, false, {600.0}, {QKA_INF}, {2.6}, {5.8}, {3.2}));
        d.push_back(iso(6346.6, 6356.0, false, {600.0}, {QKA_INF}, {2.9}, {6.8}, {3.9}));
        top = 6346.6;
    } else if (o.crust == 1) {
        d.push_back(iso(6346.6, 6371.0, false, {600.0}, {QKA_INF}, {2.6}, {5.8}, {3.2}));
        top = 6346.6;
    }
    auto upper_mantle = [&](double r0, double r1, double qmu) {
        if (o.anisotropic)
            d.push_back(ModelDomain{r0, r1, false, {qmu}, {QKA_INF}, rho_um, {0.8317, 7.2180}, {5.8582, -1.4678},
                                    {3.5908, 4.6172}, {-1.0839, 5.7176}, {3.3687, -2.4778}});
        else
            d.push_back(iso(r0, r1, false, {qmu}, {QKA_INF}, rho_um, {4.1875, 3.9382}, {2.1519, 2.3481}));
    };
    if (o.solid) {
        upper_mantle(6151.0, top, 600.0);
    } else {
        upper_mantle(6291.0, top, 600.0);                                                            // LID
        upper_mantle(6151.0, 6291.0, 80.0);                                                          // LVZ
    }
    d.push_back(iso(5971.0, 6151.0, false, {143.0}, {QKA_INF}, {7.1089, -3.8045}, {20.3926, -12.2569}, {8.9496, -4.4597}));
    d.push_back(iso(5771.0, 5971.0, false, {143.0}, {QKA_INF}, {11.2494, -8.0298}, {39.7027, -32.6166}, {22.3512, -18.5856}));
    d.push_back(iso(5701.0, 5771.0, false, {143.0}, {QKA_INF}, {5.3197, -1.4836}, {19.0957, -9.8672}, {9.9839, -4.9324}));
    const V rho_lm{7.9565, -6.4761, 5.5283, -3.0807};
    d.push_back(iso(5600.0, 5701.0, false, {312.0}, {QKA_INF}, rho_lm, {29.2766, -23.6027, 5.5242, -2.5514},
                    {22.3459, -17.2473, -2.0834, 0.9783}));
    d.push_back(iso(3630.0, 5600.0, false, {312.0}, {QKA_INF}, rho_lm, {24.9520, -40.4673, 51.4832, -26.6419},
                    {11.1671, -13.7818, 17.4575, -9.2777}));
    d.push_back(iso(3480.0, 3630.0, false, {312.0}, {QKA_INF}, rho_lm, {15.3891, -5.3181, 5.5242, -2.5514},
                    {6.9254, 1.4672, -2.0834, 0.9783}));
    d.push_back(iso(1221.5, 3480.0, !o.solid, {0.0}, {QKA_INF}, {12.5815, -1.2638, -3.6426, -5.5281},
                    {11.0487, -4.0362, 4.8023, -13.5732}, {0.0}));                                   // outer core
    d.back().vs_from_vp = o.solid;
    d.push_back(iso(0.0, 1221.5, false, {84.6}, {1327.7}, {13.0885, 0.0, -8.8381}, {11.2622, 0.0, -6.3640},
                    {3.6678, 0.0, -4.4475}));                                                        // inner core
    return d;
}

// {r_bot, rho, vp, vs} of the five CRUST2.0 layers, model_discontinuities.f90:484-828
const double CRUST20_OCEAN[5][4] = {{6367.590, 1.840, 1.920, 0.880}, {6365.330, 2.370, 3.690, 1.930}, {6364.040, 2.610, 5.090, 2.590},
                                    {6361.190, 2.900, 6.600, 3.650}, {6358.090, 3.050, 7.110, 3.910}};
const double CRUST20_CONT[5][4] = {{6370.040, 2.070, 2.420, 1.170}, {6369.210, 2.380, 3.810, 2.010}, {6356.330, 2.760, 6.130, 3.540},
                                   {6343.720, 2.880, 6.520, 3.670}, {6332.840, 3.050, 7.090, 3.930}};
const double CRUST20_GLOBAL[5][4] = {{6368.800, 1.920, 2.100, 0.980}, {6367.340, 2.370, 3.730, 1.960}, {6361.320, 2.660, 5.460, 2.920},
                                     {6355.030, 2.890, 6.570, 3.660}, {6349.180, 3.050, 7.100, 3.920}};

// ---- iasp91_sub :1836-1974 ------------------------------------------------------------------------------
std::vector<ModelDomain> iasp91() {
    std::vector<ModelDomain> d;
    const double x1 = 6251000.0 / 6371000.0, x2 = 6336000.0 / 6371000.0;      // R120, RMOHO
    const double slope = (3.3198 - 3.3713) / (x2 - x1);
    const V rho_lm{7.9565, -6.4761, 5.5283, -3.0807};
    d.push_back(iso(6351.0, 6371.0, false, {600.0}, {QKA_INF}, {2.72}, {5.8}, {3.36}));
    d.push_back(iso(6336.0, 6351.0, false, {600.0}, {QKA_INF}, {2.92}, {6.5}, {3.75}));
    d.push_back(iso(6251.0, 6336.0, false, {600.0}, {QKA_INF}, {3.3713 - slope * x1, slope}, {8.78541, -0.74953}, {6.706231, -2.248585}));
    d.push_back(iso(6161.0, 6251.0, false, {600.0}, {QKA_INF}, {2.6910, 0.6924}, {25.41389, -17.69722}, {5.75020, -1.2742}));
    d.push_back(iso(5961.0, 6161.0, false, {143.0}, {QKA_INF}, {7.1089, -3.8045}, {30.78765, -23.25415}, {15.24213, -11.08552}));
    d.push_back(iso(5711.0, 5961.0, false, {143.0}, {QKA_INF}, {5.3197, -1.4836}, {29.38896, -21.40656}, {17.70732, -13.50652}));
    d.push_back(iso(5611.0, 5711.0, false, {312.0}, {QKA_INF}, rho_lm, {25.96984, -16.93412}, {20.76890, -16.53147}));
    d.push_back(iso(3631.0, 5611.0, false, {312.0}, {QKA_INF}, rho_lm, {25.1486, -41.1538, 51.9932, -26.6083}, {12.9303, -21.2590, 27.8988, -14.1080}));
    d.push_back(iso(3482.0, 3631.0, false, {312.0}, {QKA_INF}, rho_lm, {14.49470, -1.47089}, {8.16616, -1.58206}));
    d.push_back(iso(1217.0, 3482.0, true, {0.0}, {QKA_INF}, {12.5815, -1.2638, -3.6426, -5.5281}, {10.03904, 3.75665, -13.67046}, {0.0}));
    d.push_back(iso(0.0, 1217.0, false, {84.6}, {1327.7}, {13.0885, 0.0, -8.8381}, {11.24094, 0.0, -4.09689}, {3.56454, 0.0, -3.45241}));
    return d;
}

// ---- ak135 :311-412 -------------------------------------------------------------------------------------
std::vector<ModelDomain> ak135() {
    std::vector<ModelDomain> d;
    d.push_back(iso(6351.0, 6371.0, false, {600.0}, {QKA_INF}, {2.72}, {5.8}, {3.46}));
    d.push_back(iso(6336.0, 6351.0, false, {600.0}, {QKA_INF}, {2.92}, {6.5}, {3.85}));
    d.push_back(iso(6161.0, 6336.0, false, {600.0}, {QKA_INF}, {7.1576, -3.859}, {17.4734, -9.5332}, {5.8556, -1.3825}));
    d.push_back(iso(5961.0, 6161.0, false, {143.0}, {QKA_INF}, {7.1594, -3.8608}, {30.7877, -23.2542}, {15.2181, -11.0601}));
    d.push_back(iso(5711.0, 5961.0, false, {143.0}, {QKA_INF}, {11.1204, -7.8713}, {29.389, -21.4066}, {17.7173, -13.5065}));
    d.push_back(iso(3631.0, 5711.0, false, {312.0}, {QKA_INF}, {6.8294, -1.7227, -1.1064, -0.034409}, {26.8598, -48.9644, 63.7326, -32.4155},
                    {18.0019, -43.6346, 60.4205, -29.689}));
    d.push_back(iso(3479.5, 3631.0, false, {312.0}, {QKA_INF}, {-65.8145, 386.221, -691.6551, 409.6742}, {3.4872, 55.1872, -99.0089, 58.7141},
                    {-22.9553, 164.0287, -294.2766, 174.5113}));
    d.push_back(iso(1217.5, 3479.5, true, {0.0}, {QKA_INF}, {12.592, -1.778, -1.6964, -7.3524}, {10.7738, -2.4831, 3.2584, -14.9171}, {0.0}));
    d.push_back(iso(0.0, 1217.5, false, {84.6}, {1327.7}, {13.0122, -0.0011863, -8.4449}, {11.2641, -0.090247, -5.7431}, {3.6677, 0.0049932, -4.4808}));
    return d;
}

// ---- ak135f :188-304 (Q varies with radius) -----------------------------------------------------------------
std::vector<ModelDomain> ak135f() {
    std::vector<ModelDomain> d;
    d.push_back(iso(6361.0, 6371.0, false, {599.99}, {1478.30}, {2.6}, {5.8}, {3.2}));
    d.push_back(iso(6353.0, 6361.0, false, {599.99}, {1368.02}, {2.92}, {6.8}, {3.9}));
    d.push_back(iso(6291.0, 6353.0, false, {2747.697307, -2359.725421}, {6930.368496, -5997.284866}, {199.022923, -408.226152, 212.892136},
                    {-16.800554, 50.525274, -25.691438}, {-137.314994, 285.398038, -143.603107}));
    d.push_back(iso(6251.0, 6291.0, false, {147.946500, -73.266500}, {266.958500, -86.008500}, {-8.325080, 11.977480}, {8.910012, -0.876012},
                    {6.062750, -1.592750}));
    d.push_back(iso(6161.0, 6251.0, false, {307.648222, -236.434889}, {1459.535556, -1302.515556}, {80.939817, -166.517962, 89.196989},
                    {36.839365, -41.141558, 12.026560}, {9.581677, -9.112575, 4.008853}));
    d.push_back(iso(5961.0, 6161.0, false, {528.635760, -408.763360}, {1555.736580, -1260.056380}, {25.946777, -41.561564, 18.787205},
                    {32.879415, -27.659605, 2.319408}, {6.725880, 6.913882, -9.509573}));
    d.push_back(iso(5711.0, 5961.0, false, {407.826350, -262.048331}, {762.534265, -372.539674}, {11.125709, -16.016803, 8.900728},
                    {29.313708, -21.244664, -0.086978}, {17.598241, -13.243125, -0.144963}));
    d.push_back(iso(5611.0, 5711.0, false, {-125.610200, 753.052200}, {-2797.238767, 4625.983100}, {-1.851510, 21.347947, -16.235856},
                    {37.426776, -42.812610, 14.612271}, {-36.840846, 112.512879, -72.249561}));
    d.push_back(iso(3631.0, 5611.0, false, {6.940579, 597.788236}, {301.363871, 1112.990919}, {11.655161, -23.550404, 31.096494, -15.581678},
                    {24.312100, -37.953466, 48.009009, -24.996511}, {12.135490, -18.293373, 24.249938, -12.629666}));
    d.push_back(iso(3479.5, 3631.0, false, {320.279248, -83.970345}, {719.424861, 9.016892}, {6.090533, 2.034521, -4.792620},
                    {13.185166, 2.121947, -2.292893}, {8.483334, -2.972926, 1.414776}));
    d.push_back(iso(1217.5, 3479.5, true, {0.0}, {57822.0}, {12.277066, 1.075439, -9.829445}, {10.134921, 3.305589, -13.242147}, {0.0}));
    d.push_back(iso(0.0, 1217.5, false, {85.03}, {595.258179, 166.678237}, {13.012216, -0.001140, -8.445249}, {11.264846, -0.103927, -5.687562},
                    {3.667675, 0.005479, -4.482579}));
    return d;
}

double poly(const V &c, double x) {
    double s = 0.0, xk = 1.0;
    for (double ck : c) { s += ck * xk; xk *= x; }
    return s;
}

const std::map<std::string, std::vector<ModelDomain>> &tables() {
    static const std::map<std::string, std::vector<ModelDomain>> t = [] {
        std::map<std::string, std::vector<ModelDomain>> m;
        PremOptions o;
        m["prem_iso"] = prem(o);
        o.anisotropic = true;                    m["prem_ani"] = prem(o);
        o.crust = 1;                             m["prem_ani_onecrust"] = prem(o);
        o.crust = 0;                             m["prem_ani_light"] = prem(o);
        o = PremOptions{}; o.crust = 1;          m["prem_iso_onecrust"] = prem(o);
        o.crust = 0;                             m["prem_iso_light"] = prem(o);
        o = PremOptions{}; o.solid = true;       m["prem_iso_solid"] = prem(o);
        o.crust = 0;                             m["prem_iso_solid_light"] = prem(o);
        o = PremOptions{}; o.anisotropic = true;
        o.crust20 = CRUST20_OCEAN;               m["prem_crust20_ocean"] = prem(o);
        o.crust20 = CRUST20_CONT;                m["prem_crust20_cont"] = prem(o);
        o.crust20 = CRUST20_GLOBAL;              m["prem_crust20_global"] = prem(o);
        m["iasp91"] = iasp91();
        m["ak135"] = ak135();
        m["ak135f"] = ak135f();
        return m;
    }();
    return t;
}

ExternalModel &the_external_model() {
    static ExternalModel m;
    return m;
}

std::string to_lower(std::string s) {
    for (char &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// interpolate (MESHER/interpolation.f90:118-189) on layers [i0, i1] (0-based, descending radius);
// `dx` is a default real in the reference
bool interpolate(const V &x, const V &y, int i0, int i1, bool extrapolate_constant, double xp, double &estimate) {
    const double eps = 1e-6;
    const int nd = i1 - i0 + 1;
    estimate = 0.0;
    if (!extrapolate_constant) {
        if (xp > x[i0] * (1 + eps)) return false;
        if (xp < x[i1] * (1 - eps)) return false;
    } else {
        if (xp > x[i0]) { estimate = y[i0]; return true; }
        if (xp < x[i1]) { estimate = y[i1]; return true; }
    }
    int idx = nd - 1;                               // 1-based within the range
    for (int i = 2; i <= nd - 1; i++)
        if (xp > x[i0 + i - 1]) { idx = i - 1; break; }
    const int a = i0 + idx - 1, b = a + 1;
    const float dx = (float)(x[b] - x[a]);
    if (dx != 0.0f) estimate = y[a] + (xp - x[a]) * (y[b] - y[a]) / (double)dx;
    else estimate = 0.5 * (y[b] + y[a]);
    return true;
}

}  // namespace

const std::vector<std::string> &model_names() {
    static const std::vector<std::string> n = [] {
        std::vector<std::string> v;
        for (const auto &kv : tables()) v.push_back(kv.first);
        return v;
    }();
    return n;
}

bool model_is_ani(const std::string &name) {
    if (name == "external") return external_model().anisotropic;
    return name == "prem_ani" || name == "prem_ani_onecrust" || name == "prem_ani_light" || name.rfind("prem_crust20_", 0) == 0;
}

bool model_is_anelastic(const std::string &name) {
    if (name == "external") return external_model().anelastic;
    return name != "prem_iso_solid" && name != "prem_iso_solid_light" && tables().count(name) > 0;
}

const std::vector<ModelDomain> &model_domains(const std::string &name) {
    const auto it = tables().find(name);
    if (it == tables().end()) {
        std::string all;
        for (const auto &n : model_names()) all += (all.empty() ? "" : ", ") + n;
        throw std::invalid_argument("unknown background model '" + name + "' (" + all + ", external)");
    }
    return it->second;
}

int model_domain_of(const std::string &name, double r_m, bool upper_side) {
    if (name == "external") return external_model().domain_of(r_m);
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
    if (name == "external") return external_model().evaluate(r_m, idom);
    const auto &d = model_domains(name);
    if (idom < 1 || idom > (int)d.size()) throw std::invalid_argument("idom out of range");
    const ModelDomain &m = d[idom - 1];
    const double x = (r_m / 1000.0) / 6371.0;
    ModelValues v{poly(m.rho, x) * 1000.0, poly(m.vpv, x) * 1000.0, poly(m.vsv, x) * 1000.0,
                  poly(m.vph, x) * 1000.0, poly(m.vsh, x) * 1000.0, poly(m.eta, x), poly(m.qmu, x), poly(m.qkappa, x)};
    if (m.vs_from_vp) v.vsv = v.vsh = poly(m.vpv, x) / std::sqrt(3.0) * 1000.0;
    return v;
}

// ---- external model -------------------------------------------------------------------------------------
ExternalModel parse_external_model(const std::string &text) {
    ExternalModel M;
    bool have_anel = false, have_ani = false, have_cols = false, have_units = false;
    bool in_km = true, in_depth = false, override_radius = false;
    std::map<std::string, int> col;                     // 1-based column of each quantity
    int ncolumn = 0;
    std::vector<std::vector<float>> rows;
    std::istringstream in(text);
    std::string line;
    int iline = 0;
    auto truth = [](const std::string &v) {             // Fortran list-directed logical: T, .true., true, F ...
        for (char c : v) {
            if (c == '.') continue;
            return c == 't' || c == 'T';
        }
        return false;
    };
    while (std::getline(in, line)) {
        iline++;
        size_t a = line.find_first_not_of(" \t\r");
        if (a == std::string::npos || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string keyword, value;
        ls >> keyword >> value;
        auto once = [&](bool &flag) {
            if (flag) throw std::invalid_argument("external model: parameter " + keyword + " defined twice (line " + std::to_string(iline) + ")");
            flag = true;
        };
        if (keyword == "NAME") M.name = value;
        else if (keyword == "ANELASTIC") { once(have_anel); M.anelastic = truth(value); }
        else if (keyword == "ANISOTROPIC") { once(have_ani); M.anisotropic = truth(value); }
        else if (keyword == "UNITS") { once(have_units); in_km = to_lower(value) != "m"; }
        else if (keyword == "OVERRIDE_RADIUS_CHECK") override_radius = truth(value);
        else if (keyword == "COLUMNS") {
            once(have_cols);
            std::istringstream cs(line);
            std::string tok;
            int icolumn = 0;
            while (cs >> tok) {
                const std::string t = to_lower(tok);
                if (t == "depth" || t == "radius") { col["rad"] = icolumn; in_depth = t == "depth"; }
                else if (t == "vp" || t == "vpv") col["vpv"] = icolumn;
                else if (t == "vs" || t == "vsv") col["vsv"] = icolumn;
                else if (t == "rho" || t == "qka" || t == "qmu" || t == "vph" || t == "vsh" || t == "eta") col[t] = icolumn;
                icolumn++;
            }
            ncolumn = icolumn - 1;
        } else {
            if (!have_cols) throw std::invalid_argument("external model: data before the COLUMNS line (line " + std::to_string(iline) + ")");
            std::istringstream rs(line);
            std::vector<float> r(ncolumn);
            for (int k = 0; k < ncolumn; k++) {
                double v;
                if (!(rs >> v)) throw std::invalid_argument("external model: cannot read line " + std::to_string(iline) + ": " + line);
                r[k] = (float)v;                         // layertemp is single precision
            }
            rows.push_back(r);
        }
    }
    for (const auto &kv : std::vector<std::pair<bool, const char *>>{{have_anel, "ANELASTIC"}, {have_ani, "ANISOTROPIC"},
                                                                    {have_cols, "COLUMNS"}, {have_units, "UNITS"}})
        if (!kv.first) throw std::invalid_argument(std::string("external model: parameter ") + kv.second + " is not defined");
    std::vector<std::string> need{"rad", "vpv", "vsv", "rho"};
    if (M.anelastic) { need.push_back("qka"); need.push_back("qmu"); }
    if (M.anisotropic) { need.push_back("eta"); need.push_back("vph"); need.push_back("vsh"); }
    for (const auto &n : need)
        if (!col.count(n)) throw std::invalid_argument("external model: column " + n + " is missing");
    const int nlayer = (int)rows.size();
    if (nlayer < 2) throw std::invalid_argument("external model: fewer than two layers");
    auto column = [&](const std::string &n) {
        V v(nlayer);
        for (int k = 0; k < nlayer; k++) v[k] = (double)rows[k][col[n] - 1];
        return v;
    };
    M.radius = column("rad"); M.vpv = column("vpv"); M.vsv = column("vsv"); M.rho = column("rho");
    if (M.anelastic) { M.qka = column("qka"); M.qmu = column("qmu"); }
    if (M.anisotropic) { M.vph = column("vph"); M.vsh = column("vsh"); M.eta = column("eta"); }
    double rmax = 0.0;
    for (double r : M.radius) rmax = std::fmax(rmax, r);
    if (in_km) for (double &r : M.radius) r *= 1000.0;
    else if (rmax < 10000.0 && !override_radius)
        throw std::invalid_argument("external model: radius of the model is just " + std::to_string(rmax) +
                                    " m; UNITS km, or OVERRIDE_RADIUS_CHECK true");
    if (in_depth) {
        double m = 0.0;
        for (double r : M.radius) m = std::fmax(m, r);
        for (double &r : M.radius) r = m - r;
    }
    if (M.radius[0] == 0.0) {                        // the file starts in the core: reverse
        for (V *v : {&M.radius, &M.vpv, &M.vsv, &M.rho, &M.qka, &M.qmu, &M.vph, &M.vsh, &M.eta})
            if (!v->empty()) *v = V(v->rbegin(), v->rend());
    }
    const double smallval = 1e-11;                      // smallval_dble
    if (M.radius[nlayer - 1] > smallval) M.radius[nlayer - 1] = 0.0;
    for (int k = 1; k < nlayer; k++)
        if (M.radius[k] - M.radius[k - 1] > 0.0)
            throw std::invalid_argument("external model: radius of the layers has to be monotonous (layer " + std::to_string(k + 1) + ")");

    // ---- get_ext_disc: first-order (repeated radius) and second-order (gradient step) discontinuities
    V grad_vp(nlayer - 1, 0.0), grad_vs(nlayer - 1, 0.0);
    for (int k = 0; k < nlayer - 1; k++)
        if (!(std::fabs(M.radius[k + 1] - M.radius[k]) < smallval)) {
            grad_vp[k] = (M.vpv[k + 1] - M.vpv[k]) / (M.radius[k + 1] - M.radius[k]);
            grad_vs[k] = (M.vsv[k + 1] - M.vsv[k]) / (M.radius[k + 1] - M.radius[k]);
        }
    std::vector<int> upper{1}, lower;
    for (int il = 2; il <= nlayer - 1; il++) {          // 1-based as in the reference
        const double r = M.radius[il - 1], rn = M.radius[il], rp = M.radius[il - 2];
        if (std::fabs(rn - r) < smallval) {
            lower.push_back(il);
            upper.push_back(il + 1);
        } else if ((std::fabs(grad_vp[il - 1] - grad_vp[il - 2]) >= 1e-1 || std::fabs(grad_vs[il - 1] - grad_vs[il - 2]) >= 1e-1) &&
                   r > smallval && !(std::fabs(r - rp) < smallval)) {
            lower.push_back(il);
            upper.push_back(il);
        }
    }
    if (upper.size() == 1) {                            // a blind discontinuity in the middle of the model
        upper.push_back(nlayer / 2);
        lower.push_back(nlayer / 2);
    }
    lower.push_back(nlayer);
    for (size_t k = 0; k < upper.size(); k++)
        if (upper[k] == lower[k]) upper[k] -= 1;
    M.upper_layer = upper;
    M.lower_layer = lower;
    return M;
}

ExternalModel read_external_model(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::invalid_argument("external model: file " + path + " does not exist");
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_external_model(ss.str());
}

bool ExternalModel::fluid(int idom) const {
    for (int k = upper_layer[idom - 1]; k <= lower_layer[idom - 1]; k++)
        if (vsv[k - 1] > 0.0) return false;
    return true;
}

int ExternalModel::domain_of(double r1) const {
    const int nd = ndisc();
    for (int i = 1; i <= nd - 1; i++)
        if (r1 < discont(i) && r1 > discont(i + 1)) return i;
    if (r1 < discont(nd)) return nd;
    throw std::invalid_argument("external model: have not found the domain of radius " + std::to_string(r1));
}

ModelValues ExternalModel::evaluate(double r, int idom) const {
    if (idom < 1 || idom > ndisc()) throw std::invalid_argument("external model: idom out of range");
    const int i0 = upper_layer[idom - 1] - 1, i1 = lower_layer[idom - 1] - 1;
    const bool ext = idom == 1;                          // extrapolation_constant for the first domain only
    auto at = [&](const V &y, const char *what) {
        double v;
        if (!interpolate(radius, y, i0, i1, ext, r, v))
            throw std::invalid_argument(std::string("external model: interpolation of ") + what + " not successful (layer " +
                                        std::to_string(idom) + ", radius " + std::to_string(r) + ")");
        return v;
    };
    ModelValues v;
    v.rho = at(rho, "rho");
    v.vpv = at(vpv, "vpv");
    v.vsv = at(vsv, "vsv");
    v.vph = anisotropic ? at(vph, "vph") : v.vpv;
    v.vsh = anisotropic ? at(vsh, "vsh") : v.vsv;
    v.eta = anisotropic ? at(eta, "eta") : 1.0;
    v.qmu = anelastic ? at(qmu, "qmu") : 0.0;
    v.qkappa = anelastic ? at(qka, "qka") : 0.0;
    return v;
}

void set_external_model(const ExternalModel &m) { the_external_model() = m; }

const ExternalModel &external_model() {
    if (the_external_model().radius.empty()) throw std::invalid_argument("bkgrdmodel 'external' needs a model file (--ext-model FILE.bm)");
    return the_external_model();
}

}  // namespace axisem
