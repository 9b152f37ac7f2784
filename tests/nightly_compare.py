"""End-to-end comparison with the reference's own golden seismograms
(tests/golden/nightly_ref_seismograms.npz, made by tests/golden/make_nightly_fixture.py from
TESTING/nightly/test_0{1,2,3}/ref_data/axisem.mseed).

The reference ran prem_ani (elastic) on its 50 s mesh with a 'dirac_0' source time function
(a unit-area pulse times the moment, source.f90:696-814), so its traces are, up to the
resolution of its mesh, the Green's functions of a moment step.  Here the same Earth model is
run on a synthetic mesh with a Gaussian source time function of the same area (gauss_0,
source.f90:818-831); by linearity the two agree once the reference trace is convolved with the
unit-area Gaussian.  What is restated from the post-processing (UTILS/post_processing.F90:
727-901, 922-1086): the azimuthal radiation factors of a single simulation and the rotation of
(s, phi, z) at the receiver into (E, N, Z) for a source at the north pole.

Not identical by construction, and not expected to be: the meshes differ (no inner cube and no
coarsening layers here; the source and the receivers sit on the nearest GLL point of either
mesh, which moves them by up to ~20 km).  The check is therefore waveform correlation (with a
few seconds of lag allowed for the receiver relocation) and amplitude ratio."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "nightly_ref_seismograms.npz")


def radiation_prefactor(src, lon_rad):
    """(s, phi, z) factors of compute_radiation_prefactor for a 'single' simulation."""
    one = np.ones_like(lon_rad)
    if src == "explosion":
        return one, one, one
    if src == "mtr":
        return np.cos(lon_rad), -np.sin(lon_rad), np.cos(lon_rad)
    if src == "mtp":
        return 2 * np.sin(2 * lon_rad), 2 * np.cos(2 * lon_rad), 2 * np.sin(2 * lon_rad)
    raise ValueError(src)


def to_enz(src, seis, colat_rad, lon_rad):
    """seis (nsamp, nrec, 3) = (u_s, u_phi, u_z) as the loop samples them -> (nrec, 3[E,N,Z], nsamp)."""
    fs, fp, fz = radiation_prefactor(src, lon_rad)
    us, up, uz = seis[:, :, 0] * fs, seis[:, :, 1] * fp, seis[:, :, 2] * fz
    st, ct = np.sin(colat_rad), np.cos(colat_rad)
    z = us * st + uz * ct                 # radial (up)
    n = -(us * ct - uz * st)              # north = -theta
    return np.stack([up.T, n.T, z.T], axis=1)


def compare(src, mine_enz, t_mine, t_0, decay=3.5, max_lag=6.0, window=(50.0, 1650.0), against="axisem"):
    """Per station and component: (correlation, lag [s], amplitude ratio mine/reference,
    peak amplitude of the band-limited reference trace).  against="yspec": the independent
    YSPEC solution of the explosion case instead of the reference's own traces (N and Z)."""
    z = np.load(FIXTURE)
    if against == "yspec":
        assert src == "explosion"
        nz = z["yspec_explosion_NZ"].astype(np.float64)
        ref = np.zeros((nz.shape[0], 3, nz.shape[2]))
        ref[:, 1:, :] = nz
        dt, t0 = float(z["yspec_dt"]), 0.0
    else:
        ref, dt, t0 = z[src + "_traces"].astype(np.float64), float(z[src + "_dt"]), float(z[src + "_t0"])
    tr = t0 + np.arange(ref.shape[2]) * dt
    a = decay / t_0
    tg = np.arange(-4 * t_0, 4 * t_0 + dt / 2, dt)
    g = a / np.sqrt(np.pi) * np.exp(-(a * tg) ** 2) * dt          # unit-area Gaussian of gauss_0
    w = (tr > window[0]) & (tr < window[1])
    out = np.zeros(ref.shape[:2] + (4,))
    for k in range(ref.shape[0]):
        for c in range(3):
            dc = np.convolve(ref[k, c], g, mode="same")
            best = (-2.0, 0.0, 0.0)
            for lag in np.arange(-max_lag, max_lag + 0.01, 0.5):
                m = np.interp(tr + lag, t_mine, mine_enz[k, c], left=0.0, right=0.0)
                den = np.sqrt(np.dot(m[w], m[w]) * np.dot(dc[w], dc[w]))
                cc = np.dot(m[w], dc[w]) / den if den > 0 else 0.0
                if cc > best[0]:
                    best = (cc, lag, np.dot(m[w], dc[w]) / max(np.dot(dc[w], dc[w]), 1e-300))
            out[k, c] = (best[0], best[1], best[2], np.abs(dc[w]).max())
    return out


def stations():
    z = np.load(FIXTURE)
    return list(z["names"]), z["lat"].astype(np.float64), z["lon"].astype(np.float64)
