// mapping.hpp — the element mappings of the SOLVER (reference element [-1,1]^2 -> (s, z)) and
// their partial derivatives, from the 8 control nodes of the mesher's database:
//   curved  map_spheroid / compute_partial_d_spheroid  (analytic_spheroid_mapping.f90:40-130)
//   linear  mapping_subpar / compute_partial_d_subpar   (subpar_mapping.f90:40-206, 8-node serendipity)
//   semino  map_semino / compute_partial_d_semino       (analytic_semi_mapping.f90: linear bottom, elliptic top)
//   semiso  map_semiso / compute_partial_d_semiso       (elliptic bottom, linear top)
// dispatched on eltype as analytic_mapping.f90:50-67, :582-600.
#pragma once

namespace axisem {

enum ElType { EL_CURVED = 0, EL_LINEAR = 1, EL_SEMINO = 2, EL_SEMISO = 3 };

struct MapPoint {
    double s, z, dsdxi, dzdxi, dsdeta, dzdeta;
    double jacobian() const { return dsdxi * dzdeta - dsdeta * dzdxi; }
};

// nodes[k][0] = s, nodes[k][1] = z of control node k+1 (counter-clockwise from (xi,eta) = (-1,-1))
MapPoint map_element(int eltype, const double nodes[8][2], double xi, double eta, double min_distance_dim);

}  // namespace axisem
