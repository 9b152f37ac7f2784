// receivers.hpp — the receiver side of the host set-up (SOLVER/seismograms.f90:235-639
// prepare_from_recfile_seis, SOLVER/rotations.f90:39-137): the receiver files of a run directory
// (`receivers.dat` for RECFILE_TYPE colatlon, `STATIONS` for stations), the consistency checks, and
// the rotation into the frame the solver works in (source on the north pole).  The search for the
// closest surface grid point is in precomp.cpp; the device only ever receives recfile_el(num_rec,3).
#pragma once
#include <string>
#include <vector>

namespace axisem {

struct ReceiverList {
    std::vector<std::string> name;             // recfile_0001 ... or STATION_NETWORK
    std::vector<double> colat_deg, lon_deg;    // earth-fixed frame, longitude in (0, 360]
    std::vector<std::string> redundant;        // STATIONS lines whose station + network occurred before
    size_t size() const { return name.size(); }
};

// RECFILE_TYPE colatlon: line 1 the number of receivers, then "colatitude longitude" [deg] per line
ReceiverList read_receivers_dat(const std::string &path);
// RECFILE_TYPE stations: "name network latitude longitude elevation burial" per line; repeated
// station + network pairs are dropped (and listed in `redundant`), longitudes <= 0 get + 360
ReceiverList read_stations(const std::string &path);
// the reference's stops: negative or > 360.001 longitudes, colatitudes outside [0, 180.001]
void check_receiver_coordinates(const ReceiverList &r);

// rotate_receivers_recfile: receiver coordinates in the frame with the source at the north pole
// (rotation matrix of Nissen-Meyer, Dahlen & Fournier 2007, def_rot_matrix); angles in degrees
void rotate_receivers(double srccolat_rad, double srclon_rad, std::vector<double> &colat_deg, std::vector<double> &lon_deg);

// receiver_names.dat ("name colat lon" as read) and, for a rotated source, receiver_rotated.dat
void write_receiver_names(const std::string &path, const ReceiverList &r);
void write_receiver_rotated(const std::string &path, const std::vector<double> &colat_deg, const std::vector<double> &lon_deg);

// ---- what the two host tools do with the above --------------------------------------------------------
struct ReceiverSetup {
    std::string receivers_file, stations_file;       // one of them (RECFILE_TYPE colatlon | stations)
    double src_lat_deg = 90.0, src_lon_deg = 0.0;    // SOURCE_LAT, SOURCE_LON of inparam_source
    bool given() const { return !receivers_file.empty() || !stations_file.empty(); }
    bool rot_src() const { return src_lat_deg != 90.0 || src_lon_deg != 0.0; }
};
// reads and checks the list, rotates it when the source is not on the north pole, writes
// PREFIX.receiver_names.dat (as read) and PREFIX.receiver_rotated.dat (the solver's frame: what
// axisem_b200_postproc --stations takes); returns the coordinates in the solver's frame
ReceiverList prepare_receivers(const ReceiverSetup &s, const std::string &prefix, std::vector<double> &colat_deg,
                               std::vector<double> &lon_deg);
// PREFIX.receiver_pts.dat: "colatitude of the grid point taken, longitude, rank" per receiver of the list
// (seismograms.f90:540-546); loc2globrec / recfile_th per rank as the pre-computation stored them
void write_receiver_pts(const std::string &path, const std::vector<std::vector<int>> &loc2globrec,
                        const std::vector<std::vector<double>> &recfile_th, const std::vector<double> &lon_deg);

}  // namespace axisem
