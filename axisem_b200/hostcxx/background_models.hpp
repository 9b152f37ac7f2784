// background_models.hpp — the 1-D models the SOLVER evaluates in get_model: every internal model of
// SOLVER/background_models.F90 (`velocity`, :69-113) as a table of per-domain polynomials, and the
// `external` model (a tabulated .bm file: read_ext_model :2082-2421, get_ext_disc :2618-2784,
// arbitr_sub_solar :1983-2078, MESHER/interpolation.f90:118-189).
// Domains are numbered from the surface inwards as in the reference (idom = 1 at the surface); their
// radii are the reference's MESHER/model_discontinuities.f90.  prem_ani is pinned against the
// reference's own tabulation (TESTING/TEST04_anelastic_anisotropic/model.bm,
// tests/test_reference_fixtures.py); the external reader against the same table.
#pragma once
#include <string>
#include <vector>

namespace axisem {

struct ModelDomain {
    double r_bot_km, r_top_km;
    bool fluid;
    std::vector<double> qmu, qkappa;                      // polynomial coefficients in x = r / 6371 km
    std::vector<double> rho, vpv, vsv, vph, vsh, eta;
    bool vs_from_vp = false;                               // prem_*_solid*: v_s = v_p / sqrt(3) in the outer core
};

struct ModelValues {
    double rho, vpv, vsv, vph, vsh, eta, qmu, qkappa;    // SI units (kg/m^3, m/s)
};

// every bkgrdmodel of the reference except `external`
const std::vector<std::string> &model_names();
bool model_is_ani(const std::string &bkgrdmodel);        // background_models.F90:118-141
bool model_is_anelastic(const std::string &bkgrdmodel);  // :146-181

const std::vector<ModelDomain> &model_domains(const std::string &bkgrdmodel);

// idom (1-based, from the surface) of radius r [m]; on a discontinuity `upper_side` picks the
// domain above it
int model_domain_of(const std::string &bkgrdmodel, double r_m, bool upper_side);

ModelValues model_evaluate(const std::string &bkgrdmodel, double r_m, int idom);

// ---- bkgrdmodel = 'external' ----------------------------------------------------------------------
struct ExternalModel {
    std::string name = "external_model";
    bool anelastic = false, anisotropic = false;
    // layers from the surface to the centre, SI units (the file's values pass through single precision
    // as in the reference: `real(kind=sp) :: layertemp`)
    std::vector<double> radius, rho, vpv, vsv, qka, qmu, vph, vsh, eta;
    // domains between discontinuities: 1-based layer ranges [upper, lower] (get_ext_disc)
    std::vector<int> upper_layer, lower_layer;
    int ndisc() const { return (int)upper_layer.size(); }
    double discont(int idom) const { return radius[upper_layer[idom - 1] - 1]; }
    bool fluid(int idom) const;
    int domain_of(double r_m) const;                      // get_model.F90:120-146 on the model's own discont
    ModelValues evaluate(double r_m, int idom) const;     // arbitr_sub_solar; throws where the reference stops
};

// parses the text of a .bm file (throws std::invalid_argument with the reference's error conditions)
ExternalModel parse_external_model(const std::string &text);
ExternalModel read_external_model(const std::string &path);

// makes `model` the one the name "external" refers to in model_domain_of / model_evaluate
void set_external_model(const ExternalModel &model);
const ExternalModel &external_model();

}  // namespace axisem
