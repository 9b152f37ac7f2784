// spectral.hpp — what data_spec holds (SOLVER/data_spec.f90:31-39) for one npol: GLL nodes and
// weights (eta, wt), the Gauss-Lobatto-Jacobi(0,1) nodes and weights of the axial elements
// (xi_k, wt_axial_k), and the derivative matrices G0, G1, G1T, G2, G2T in realkind.
// Counterpart of MESHER/gllmeshgen.f90:61-94 + MESHER/splib.f90 (zelegl, zemngl2, get_welegl,
// get_welegl_axial, hn_jprime, lag_interp_deriv_wgl); the formulas, not the routines.
#pragma once
#include <vector>

namespace axisem {

struct SpectralBasis {
    int npol = 0;
    std::vector<double> eta, wt, xi_k, wt_axial_k;      // (0:npol)
    std::vector<double> G1_dp, G2_dp;                   // [j + (npol+1)*i] = l_j'(x_i): Fortran G(j,i)
    std::vector<float> G0, G1, G1T, G2, G2T;            // as stored in the mesh database
};

SpectralBasis spectral_basis(int npol);

}  // namespace axisem
