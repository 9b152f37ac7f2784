// background_models.hpp — the polynomial Earth models the SOLVER evaluates in get_model
// (SOLVER/background_models.F90: prem_sub :417-529, prem_ani_sub :534-674), table-driven.
// Domains are numbered from the surface inwards as in the reference (idom = 1 upper crust ...
// 12 inner core).  Pinned against the reference's own tabulation of prem_ani
// (TESTING/TEST04_anelastic_anisotropic/model.bm, tests/test_reference_fixtures.py).
#pragma once
#include <string>
#include <vector>

namespace axisem {

struct ModelDomain {
    double r_bot_km, r_top_km;
    bool fluid;
    double qmu, qkappa;
    std::vector<double> rho, vpv, vsv, vph, vsh, eta;   // polynomial coefficients in x = r / 6371 km
};

struct ModelValues {
    double rho, vpv, vsv, vph, vsh, eta, qmu, qkappa;    // SI units (kg/m^3, m/s)
};

// "prem_iso" or "prem_ani"
const std::vector<ModelDomain> &model_domains(const std::string &bkgrdmodel);

// idom (1-based, from the surface) of radius r [m]; on a discontinuity `upper_side` picks the
// domain above it
int model_domain_of(const std::string &bkgrdmodel, double r_m, bool upper_side);

ModelValues model_evaluate(const std::string &bkgrdmodel, double r_m, int idom);

}  // namespace axisem
