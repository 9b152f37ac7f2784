// axisem_b200_meshdb2axbp — MESHER output -> module variables.
//   axisem_b200_meshdb2axbp meshdb.dat0003 3 rank3_mesh.axbp
// reads the mesher's database of one rank (meshdb.cpp) and writes the mesh-level module
// variables (data_mesh, data_spec, data_time, data_comm) as an AXBPROB1 container.
#include <cstdio>
#include <cstdlib>

#include "meshdb.hpp"

int main(int argc, char **argv) {
    if (argc != 4) {
        std::fprintf(stderr, "usage: axisem_b200_meshdb2axbp meshdb.datNNNN mynum out.axbp\n");
        return 2;
    }
    try {
        const axisem::Modules m = axisem::read_meshdb(argv[1], std::atoi(argv[2]));
        axisem::write_container(m, argv[3]);
        std::printf("%s: nproc %d, npol %d, %d solid + %d fluid elements, %d S/F boundary elements, %zu variables\n",
                    argv[1], m.int_of("data_proc%nproc"), m.int_of("data_mesh%npol"), m.int_of("data_mesh%nel_solid"),
                    m.int_of("data_mesh%nel_fluid"), m.int_of("data_mesh%nel_bdry"), m.size());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
    return 0;
}
