// rundir.hpp — the files a run of the reference leaves behind when USE_NETCDF is false, which is what
// its own post-processing (SOLVER/UTILS/post_processing.F90) reads: `simulation.info`
// (parameters.F90:1410-1465, formats 21-25; read back at post_processing.F90:614-647),
// `Data/receiver_names.dat`, `Data/receiver_pts.dat` (seismograms.f90:276-288, 540-546) and one
// `Data/<receiver>_disp.dat` per receiver (compute_recfile_seis_bare, seismograms.f90:742-778: u_s, u_z for a
// monopole, u_s, u_phi, u_z otherwise, one line per seismogram sample).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace axisem {

struct SimulationInfo {
    std::string bkgrdmodel, src_type1, src_type2, stf_type, simtype = "single", rec_comp = "cyl", rec_file_type = "colatlon";
    double deltat = 0, period = 0, src_depth_km = 0, srccolat = 0, srclon = 0, magnitude = 0;
    int niter = 0, num_rec_tot = 0, nseismo = 0;
    double seis_dt = 0;
    int nstrain = 0;
    double strain_dt = 0;
    int nsnap = 0;
    double snap_dt = 0;
    int ibeg = 0, iend = 4;
    double shift_fact = 0;
    int ishift_deltat = 0, ishift_seisdt = 0, ishift_straindt = 0;
    double dtheta_rec = 0;
    bool use_netcdf = false;
    int nelem = 0, nel_fluid = 0, nproc = 1;
};

void write_simulation_info(const std::string &path, const SimulationInfo &s);
// list-directed reads of the first item of each line, in the order of post_processing.F90:614-647
SimulationInfo read_simulation_info(const std::string &path);

// seis: (nseis, nrec, 3) as OutputSink::seismograms delivers recdumpvar(3, num_rec, nseis), receivers in
// the order of `names`
void write_disp_files(const std::string &data_dir, const std::vector<std::string> &names, bool monopole,
                      int nseis, const std::vector<float> &seis);
// the way back (post_processing.F90:344-357): (nseis, nrec, 3), u_phi = 0 for a monopole
std::vector<float> read_disp_files(const std::string &data_dir, const std::vector<std::string> &names, bool monopole, int nseis);

// XDMF snapshots of one rank as the reference leaves them in Data/ (dump_xdmf_grid, meshes_io.F90:400-437;
// glob_snapshot_xdmf formats 733-735 and finish_xdmf_xml, wavefields_io.f90:195-330): big-endian
// xdmf_points_NNNN.dat, xdmf_grid_NNNN.dat, xdmf_snap_{s,p,z,trace,curlip}_NNNN.dat (no p file for a
// monopole), xdmf_meshonly_NNNN.xdmf, xdmf_xml_NNNN.xdmf.  fields: (5, nsnap, npoint_plot) as
// OutputSink::xdmf delivers them; times[k]: time of snapshot k
void write_xdmf_files(const std::string &data_dir, int rank, int npoint_plot, int nelem_plot, const float *points,
                      const int32_t *grid, const float *fields, int nsnap, const std::vector<double> &times, bool monopole);

void make_directory(const std::string &path);      // mkdir -p of one level; no error if it exists

}  // namespace axisem
