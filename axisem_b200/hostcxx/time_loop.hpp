// time_loop.hpp — the host side of the seam `call time_loop` (SOLVER/main.f90:92).
//
// `time_loop` below is what a maintainer's replacement of SOLVER/time_evol_wave.F90:231-245
// does, written in C++ because this image has no Fortran compiler (INTEGRATION.md shows the
// same sequence with iso_c_binding): hand the module arrays to the device library once,
// advance the device-resident state in chunks, and pass the receiver / wavefield buffers to
// the output layer where the reference calls nc_dump_rec / nc_dump_field_* from dump_stuff
// (time_evol_wave.F90:1104-1251).  Same control flow and error behaviour as the reference:
// progress line every 100 steps (runtime_info :1009-1018), seismogram check-point every
// check_it = niter/20 steps (parameters.F90:932, :1142), wavefield buffers flushed every
// nc_dumpbuffersize snapshots (parameters.F90:418), "DISPLACEMENTS BLEW UP" ends the run.
//
// One process can drive every theta-slice of a box (one handle per GPU, peers wired with
// direct pointers over NVLink): pass one Modules per rank.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "modules.hpp"

namespace axisem {

// where the reference's NetCDF writers sit (nc_routines.F90:530-540, 248-275)
class OutputSink {
public:
    virtual ~OutputSink() {}
    // recdumpvar(3, num_rec, first+1 : first+n) of rank `rank`
    virtual void seismograms(int rank, int num_rec, int first, int n, const float *recdumpvar) = 0;
    // oneddumpvar(npoints, first+1 : first+n, nvars); nvars = nvar/2 of nc_routines.F90:943-1050
    // (3 for displ_only, 6 (4) for strain_only, 9 (6) for fullfields)
    virtual void snapshots(int rank, size_t npoints, int nvars, int first, int n, const float *oneddumpvar) = 0;
    // xdmf snapshots (glob_snapshot_xdmf, wavefields_io.f90:195-199): (npoint_plot, n, 5) = u_s, u_p,
    // u_z, straintrace, curlinplane of all n snapshots of the run
    virtual void xdmf(int rank, size_t npoint_plot, int n, const float *fields) { (void)rank; (void)npoint_plot; (void)n; (void)fields; }
    // dump_energy: (4, n) sums of this rank for iter 0..n-1 (time_evol_wave.F90:1424-1526; the
    // writer applies psum over ranks and two*pi)
    virtual void energy(int /*rank*/, int /*n*/, const float * /*sums*/) {}
};

struct TimeLoopOptions {
    int nsteps = -1;              // -1: niter of data_time
    int check_iter = 100;         // runtime_info progress cadence
    int nc_dumpbuffersize = 128;  // snapshots buffered before they go to the sink
    int ndevices = 1;             // rank r runs on device r % ndevices
    bool verbose = true;          // lpr
    FILE *log = stdout;
};

struct TimeLoopResult {
    int iter = 0, nseismo = 0, nstrain = 0;
    int64_t gpu_launches = 0;
    double seconds = 0.0;         // wall clock of the stepping (not the set-up)
};

// Runs the whole loop for the given ranks (all slices of the run that live in this process).
// Throws SolverError with the library's message on any failure.
TimeLoopResult time_loop(const std::vector<Modules> &ranks, const TimeLoopOptions &opt, OutputSink *sink);

}  // namespace axisem
