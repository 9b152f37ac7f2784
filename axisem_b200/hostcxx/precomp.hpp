// precomp.hpp — from the mesher's database to the module variables the time loop works on,
// natively: what the SOLVER does between `read_db` and `call time_loop`.
//
//   element geometry          analytic_mapping.f90 (mapping.hpp), def_grid.f90 (coordinates)
//   background model          get_model.F90:155-186 with background_models.hpp
//   mass matrices             def_precomp_terms.f90:596-835 (def_mass_matrix_k, assembled with the
//                             halo partners as pdistsum does, inverted; dipole factor :773)
//   solid stiffness planes    def_precomp_terms.f90:1166-2332 (mono / di / quad, TI via c_ijkl_ani,
//                             axial M0_w* vectors, anelastic Y / V planes and their cg4 samples)
//   fluid stiffness planes    def_precomp_terms.f90:2336-2470
//   S/F boundary terms        def_precomp_terms.f90:2474-2712
//   pointwise derivatives     def_precomp_terms.f90:148-304
//   attenuation               attenuation.f90:882-1091 (time-step factors, delta / unrelaxed moduli,
//                             coarse-grained weights; the SLS fit itself is an input)
//   source                    source.f90:206-233, 454-476, 587-660, 921-1226 (on-axis point source)
//   receivers                 seismograms.f90:235-639 (nearest surface GLL point)
//   wavefield-dump point set  meshes_io.F90:489-640
//   xdmf plot points and grid meshes_io.F90:110-437 (dump_xdmf_grid)
// Every array is stored under the reference's `<module>%<variable>` name in Fortran memory order,
// i.e. exactly what axisem_b200/hostcxx/time_loop.cpp hands to the C ABI.
#pragma once
#include <string>
#include <vector>

#include "modules.hpp"

namespace axisem {

struct AttenuationOptions {
    bool coarse_grained = true, do_corr_lowq = true;
    // a 5-SLS log-spaced fit for 1 mHz - 1 Hz (the fit itself — simulated annealing with an
    // unseeded RNG, attenuation.f90:1183-1339 — is not reproducible; w_j, y_j are inputs)
    std::vector<double> w_j = {2 * 3.14159265358979323846 * 0.0015, 2 * 3.14159265358979323846 * 0.0090,
                               2 * 3.14159265358979323846 * 0.052, 2 * 3.14159265358979323846 * 0.29,
                               2 * 3.14159265358979323846 * 1.55};
    std::vector<double> y_j = {1.53, 1.16, 1.21, 1.09, 1.70};
    double f_min = 0.001, f_max = 1.0, w_0 = 1.0;
};

struct PrecompOptions {
    std::string model;                  // "" = the database's bkgrdmodel (prem_iso | prem_ani)
    std::string src_type2 = "explosion";
    double src_depth = 100.0e3, magnitude = 1.0e20, t_0 = 50.0, decay = 3.5, shift_fact = 1.5;
    std::string stf_type = "gauss_0";     // gauss_0|gauss_1|gauss_2|errorf|dirac_0|dirac_1|quheavi (source.f90:152-171)
    std::string discrete_choice = "gaussi"; // delta_src's approximation of the Dirac (parameters.F90:995-999)
    double shift_seconds = -1.0;          // >= 0: shift_fact in seconds as the caller fixed it; else from shift_fact * t_0
    std::string time_scheme = "newmark2";
    int niter = 100, seis_it = 1, strain_it = 0;
    double deltat = 0.0;                // 0 = the mesher's (times 1.5 for the symplectic schemes)
    bool attenuation = false;
    AttenuationOptions att;
    bool dump_wavefields = false;
    bool dump_energy = false;
    std::vector<double> rec_colat_deg;  // receivers on the surface
    // xdmf snapshots (SAVE_SNAPSHOTS with SNAPSHOTS_FORMAT xdmf): snap_it = floor(SNAPSHOT_DT / deltat),
    // XDMF_GLL_I / _J, XDMF_RMIN / RMAX [m], XDMF_COLAT_MIN / MAX [rad] (parameters.F90:424-431, 572-594, 944)
    int snap_it = 0;
    std::vector<int> xdmf_gll_i = {0, 2, 4}, xdmf_gll_j = {0, 2, 4};
    double xdmf_rmin = 0.0, xdmf_rmax = 7.0e6, xdmf_thetamin = 0.0, xdmf_thetamax = 3.14159265358979323846;
};

// The reference's fit of n_sls standard linear solids to constant Q over [f_min, f_max] Hz (invert_linear_solids,
// attenuation.f90:1183-1339, with the defaults of inparam_advanced), seeded; fills A.w_j, A.y_j (for Q = 1), A.f_min,
// A.f_max and returns the frequency-weighted log-l2 misfit
double fit_linear_solids(AttenuationOptions &A, int n_sls, double f_min, double f_max, uint64_t seed = 0, int max_it = 100000);
std::vector<double> q_linear_solid(const std::vector<double> &y_j, const std::vector<double> &w_j, const std::vector<double> &w);

// loc2globrec / recfile_th of every rank after precompute (for receiver_pts.dat)
std::vector<std::vector<int>> receiver_indices(const std::vector<Modules> &ranks);
std::vector<std::vector<double>> receiver_colatitudes(const std::vector<Modules> &ranks);

// `ranks`: read_meshdb results of all ranks of the run, in rank order (mass matrices are
// assembled across the cuts).  On return every Modules also holds the time-loop inputs.
void precompute(std::vector<Modules> &ranks, const PrecompOptions &opt);

// the reference's self-checks on what was computed (def_grid.f90:1188 "mass = volume",
// def_precomp_terms.f90:2743 "S/F boundary term = 2 per boundary"), summed over ranks
struct PrecompChecks { double solid_volume, fluid_volume, sphere_volume, hollow_volume, bdry_sum; int n_sf_boundaries; };
PrecompChecks precompute_checks(const std::vector<Modules> &ranks);

}  // namespace axisem
