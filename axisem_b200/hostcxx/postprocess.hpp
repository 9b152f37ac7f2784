// postprocess.hpp — what follows the seam on the way to seismograms
// (SOLVER/UTILS/post_processing.F90): the azimuthal radiation factors of each simulation
// (compute_radiation_prefactor :727-901), the sum over simulations (sum_individual_wavefields
// :905-918), the rotation of (s, phi, z) at the receiver into the requested component system
// (rotate_receiver_comp :922-1086) and the convolution with a Gaussian source time function
// (convolve_with_stf).  Source at the north pole (the solver's own frame); the general source
// location adds the rotation of :745-766 in front, which is not part of this file yet.
// The formulas are pinned by the reference's golden seismograms (tests/test_nightly_reference.py
// through their Python restatement, tests/test_host_postprocess.py through this one).
#pragma once
#include <string>
#include <vector>

namespace axisem {

// Mij = (Mrr, Mtt, Mpp, Mrt, Mrp, Mtp) [N m] of the event; `magnitude` of the simulation.
// Returns the factors of (u_s, u_phi, u_z) for a simulation of type src_type2 at longitude lon.
void radiation_prefactor(const std::string &src_type2, const double Mij[6], double magnitude, double lon_rad,
                         double out[3]);

// Mij of a 'single' simulation: the moment tensor the source type stands for, times amplitude
void single_simulation_moment(const std::string &src_type2, double amplitude, double Mij[6]);

// seis: (nsamp, 3) = (u_s, u_phi, u_z) already multiplied by the radiation factors and summed.
// comp_sys: "enz" (east, north, up), "sph" (r, theta, phi), "cyl" (s, phi, z)
void rotate_receiver_comp(const std::string &comp_sys, double colat_rad, size_t nsamp, const float *seis_spz,
                          float *out3);

// unit-area Gaussian of gauss_0 (source.f90:818-831): a = decay / t_0
void convolve_gauss(std::vector<float> &trace, double dt, double t_0, double decay);

}  // namespace axisem
