// meshdb.hpp — reader of the MESHER's per-rank mesh database `meshdb.datNNNN`.
//
// Native counterpart of SOLVER/data_mesh.f90:190-322 (read_mesh_basics / _advanced / _axel)
// and SOLVER/get_mesh.f90:101-383 (read_db): Fortran sequential-unformatted records in the
// order MESHER/pdb.f90:2205-2382 writes them.  The result is a `Modules` holding what those
// routines leave in the Fortran modules (data_mesh, data_spec, data_time, data_comm), plus
// what def_grid derives from it for the time loop (SOLVER/def_grid.f90:59-77 axis flags,
// :95-180 glob2el_* / num_comm_gll_*).  Geometry-dependent pre-computation
// (def_precomp_terms, get_model) is not part of this reader.
#pragma once
#include <string>

#include "modules.hpp"

namespace axisem {

Modules read_meshdb(const std::string &path, int mynum);

// write a Modules as an AXBPROB1 container (the format Modules::read takes)
void write_container(const Modules &m, const std::string &path);

}  // namespace axisem
