// axisem_b200_meshdb2axbp — MESHER output -> module variables.
//   axisem_b200_meshdb2axbp meshdb.dat0003 3 rank3_mesh.axbp
// reads the mesher's database of one rank (meshdb.cpp) and writes the mesh-level module
// variables (data_mesh, data_spec, data_time, data_comm) as an AXBPROB1 container.
#include <cstdio>
#include <cstdlib>

#include <cmath>

#include "meshdb.hpp"
#include "spectral.hpp"

int main(int argc, char **argv) {
    if (argc != 4) {
        std::fprintf(stderr, "usage: axisem_b200_meshdb2axbp meshdb.datNNNN mynum out.axbp\n");
        return 2;
    }
    try {
        const axisem::Modules m = axisem::read_meshdb(argv[1], std::atoi(argv[2]));
        axisem::write_container(m, argv[3]);
        // the database's spectral arrays against the native basis (the SOLVER trusts the mesher here)
        const int npol = m.int_of("data_mesh%npol");
        const axisem::SpectralBasis b = axisem::spectral_basis(npol);
        double dev = 0.0;
        const double *eta = m.d("data_spec%eta"), *xi = m.d("data_spec%xi_k");
        const float *G1 = m.f("data_spec%G1"), *G2 = m.f("data_spec%G2");
        for (int k = 0; k <= npol; k++) {
            dev = std::fmax(dev, std::fabs(eta[k] - b.eta[k]));
            dev = std::fmax(dev, std::fabs(xi[k] - b.xi_k[k]));
        }
        for (size_t k = 0; k < b.G1.size(); k++) {
            dev = std::fmax(dev, std::fabs((double)G1[k] - (double)b.G1[k]));
            dev = std::fmax(dev, std::fabs((double)G2[k] - (double)b.G2[k]));
        }
        std::printf("spectral arrays of the database vs native basis: max deviation %.3e\n", dev);
        if (dev > 1e-5) throw axisem::SolverError("the database's GLL/GLJ arrays do not belong to npol = " + std::to_string(npol));
        std::printf("%s: nproc %d, npol %d, %d solid + %d fluid elements, %d S/F boundary elements, %zu variables\n",
                    argv[1], m.int_of("data_proc%nproc"), m.int_of("data_mesh%npol"), m.int_of("data_mesh%nel_solid"),
                    m.int_of("data_mesh%nel_fluid"), m.int_of("data_mesh%nel_bdry"), m.size());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
