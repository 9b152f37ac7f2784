// axisem_b200_hosttool — prints what the native pre-processing pieces compute, for the tests:
//   hosttool spectral NPOL                    eta, wt, xi_k, wt_axial_k, G0, G1, G1T, G2, G2T
//   hosttool model NAME                       radii + side (u/l) on stdin -> rho vpv vsv vph vsh eta qka qmu idom
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "background_models.hpp"
#include "spectral.hpp"

template <class T>
static void line(const char *name, const std::vector<T> &v) {
    std::printf("%s", name);
    for (T x : v) std::printf(" %.17g", (double)x);
    std::printf("\n");
}

int main(int argc, char **argv) {
    try {
        if (argc == 3 && !std::strcmp(argv[1], "spectral")) {
            const axisem::SpectralBasis b = axisem::spectral_basis(std::atoi(argv[2]));
            line("eta", b.eta); line("wt", b.wt); line("xi_k", b.xi_k); line("wt_axial_k", b.wt_axial_k);
            line("G0", b.G0); line("G1", b.G1); line("G1T", b.G1T); line("G2", b.G2); line("G2T", b.G2T);
            return 0;
        }
        if (argc == 3 && !std::strcmp(argv[1], "model")) {
            double r;
            char side;
            while (std::scanf("%lf %c", &r, &side) == 2) {
                const int idom = axisem::model_domain_of(argv[2], r, side == 'u');
                const axisem::ModelValues v = axisem::model_evaluate(argv[2], r, idom);
                std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d\n", v.rho, v.vpv, v.vsv, v.vph, v.vsh,
                            v.eta, v.qkappa, v.qmu, idom);
            }
            return 0;
        }
        std::fprintf(stderr, "usage: axisem_b200_hosttool spectral NPOL | model NAME < radii\n");
        return 2;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return 1;
    }
}
