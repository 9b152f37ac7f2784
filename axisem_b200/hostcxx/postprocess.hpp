// postprocess.hpp — what follows the seam on the way to seismograms
// (SOLVER/UTILS/post_processing.F90), restated for the files axisem_b200_solver writes:
//   * receiver_location            :187-232  receiver coordinates in the earth-fixed frame from
//                                            those in the solver's frame (source at the pole)
//   * moment_from_cmtsolution      :774-793  Mij of simtype 'moment' (dyn cm -> N m)
//   * single_simulation_moment     :795-836  Mij of simtype 'single'
//   * radiation_prefactor          :838-900  azimuthal factors of each of the (up to 4) runs
//   * sum_individual_wavefields    :905-918
//   * convolve_with_stf            :1014-1084 (gauss_0, gauss_1; causal, shifted by 1.5 t_0)
//   * rotate_receiver_comp         :922-1009 all five component systems, source anywhere
// Component order of the results is the reference's: enz -> (N, E, Z), sph -> (theta, phi, r),
// cyl -> (s, phi, z), xyz, src -> (theta', phi', r') of the source-centred frame
// (reccomp of post_processing.F90:272-296).
#pragma once
#include <string>
#include <vector>

namespace axisem {

constexpr double POST_DECAY = 3.5, POST_SHIFT_FACT1 = 1.5;     // post_processing.F90:97-98

struct SourceLocation { double colat = 0.0, lon = 0.0; };       // radians

// Mij = (Mrr, Mtt, Mpp, Mrt, Mrp, Mtp) [N m]
void single_simulation_moment(const std::string &src_type2, double amplitude, double Mij[6]);
// the seven header lines and six "name: value" lines of a CMTSOLUTION file; values in dyn cm
void moment_from_cmtsolution(const std::string &path, double Mij[6]);

// factors of (u_s, u_phi, u_z) of a run of type src_type2 with source magnitude `magnitude`
// at receiver longitude lon (solver frame)
void radiation_prefactor(const std::string &src_type2, const double Mij[6], double magnitude, double lon_rad,
                         double out[3]);

// field_sum(:, c) += prefactor(c) * field_in(:, c); fields are (nsamp, 3) = s, phi, z interleaved per sample
void sum_individual_wavefields(std::vector<float> &field_sum, const float *field_in, size_t nsamp,
                               const double prefactor[3]);

// (colat, lon) of a receiver in the earth-fixed frame, given its coordinates in the solver frame
void receiver_location(const SourceLocation &src, double colat_rot, double lon_rot, double &colat_orig,
                       double &lon_orig);

// seis (nsamp, 3) in place
void rotate_receiver_comp(const std::string &rec_comp_sys, const SourceLocation &src, double th_rot,
                          double ph_rot, double th_orig, double ph_orig, size_t nsamp, float *seis);

// the reference's causal convolution (result delayed by 1.5 t_0): seis_fil(i) = pi * sum_j seis(i-j) stf(j dt) dt
void convolve_with_stf(double t_0, double dt, const std::string &stf, size_t nsamp, const float *seis, float *seis_fil);

// zero-phase unit-area Gaussian of gauss_0 (source.f90:818-831), a = decay / t_0: what the
// comparison with the nightly traces uses (tests/nightly_compare.py)
void convolve_gauss(std::vector<float> &trace, double dt, double t_0, double decay);

}  // namespace axisem
