#!/usr/bin/env python
"""The reference's own golden seismograms -> tests/golden/nightly_ref_seismograms.npz.

TESTING/nightly/test_0{1,2,3}/ref_data/axisem.mseed are the traces the reference's nightly
regression compares every new build against: PREM (prem_ani, elastic), 50 s mesh, source at
the north pole in 100 km depth with 1e20 Nm, 'dirac_0' source time function, explosion
(monopole) / mtr (dipole) / mtp (quadrupole), 20 stations x (E, N, Z), 1800 s of
displacement at the solver's own sampling (0.859 s), time axis starting at the negative source
shift.  miniSEED: 4096-byte records, blockette 1000, encoding 5 (IEEE float64, big endian).
Run in the build container (needs /root/reference); the .npz travels."""
import os
import struct

import numpy as np

ROOT = "/root/reference/TESTING/nightly"


def read_mseed(path):
    data = open(path, "rb").read()
    recs, pos, reclen = {}, 0, 4096
    while pos + 48 <= len(data):
        h = data[pos:pos + 48]
        sta, cha = h[8:13].decode().strip(), h[15:18].decode().strip()
        yr, doy, hh, mm, ss, _, frac = struct.unpack(">HHBBBBH", h[20:30])
        (nsamp,) = struct.unpack(">H", h[30:32])
        rfac, rmul = struct.unpack(">hh", h[32:36])
        nblk = h[39]
        dataoff, b = struct.unpack(">HH", h[44:48])
        enc = None
        for _ in range(nblk):
            btype, bnext = struct.unpack(">HH", data[pos + b:pos + b + 4])
            if btype == 1000:
                enc, reclen = data[pos + b + 4], 2 ** data[pos + b + 6]
            if bnext == 0:
                break
            b = bnext
        assert enc == 5, enc
        if rfac > 0 and rmul > 0:
            sr = rfac * rmul
        elif rfac > 0 and rmul < 0:
            sr = -rfac / rmul
        elif rfac < 0 and rmul > 0:
            sr = -rmul / rfac
        else:
            sr = 1.0 / (rfac * rmul)
        d = np.frombuffer(data[pos + dataoff:pos + dataoff + 8 * nsamp], dtype=">f8")
        t0 = ((doy - 1) * 86400 + hh * 3600 + mm * 60 + ss + frac * 1e-4) - (0 if yr == 1970 else 365 * 86400)
        r = recs.setdefault((sta, cha), {"sr": sr, "t0": t0, "d": []})
        r["d"].append(d)
        pos += reclen
    return {k: (v["sr"], v["t0"], np.concatenate(v["d"]).astype(np.float64)) for k, v in recs.items()}


out = {}
for test, src in (("test_01", "explosion"), ("test_02", "mtr"), ("test_03", "mtp")):
    st = [l.split() for l in open(os.path.join(ROOT, test, "STATIONS")) if l.strip()]
    names = [s[0] for s in st]
    tr = read_mseed(os.path.join(ROOT, test, "ref_data", "axisem.mseed"))
    sr, t0, d0 = tr[(names[0], "Z")]
    arr = np.zeros((len(names), 3, d0.size), dtype=np.float32)
    for i, n in enumerate(names):
        for c, comp in enumerate("ENZ"):
            s, t, d = tr[(n, comp)]
            assert abs(s - sr) < 1e-12 and abs(t - t0) < 1e-9 and d.size == d0.size
            arr[i, c] = d
    out[src + "_traces"] = arr                                  # (station, E/N/Z, sample) [m]
    out[src + "_dt"] = np.float64(1.0 / sr)
    out[src + "_t0"] = np.float64(t0)                           # time of sample 0 relative to the origin time
    out["lat"] = np.array([float(s[2]) for s in st])
    out["lon"] = np.array([float(s[3]) for s in st])
    out["names"] = np.array(names)
    print(test, src, arr.shape, 1.0 / sr, t0, np.abs(arr).max())
# The independent solution the reference ships next to its own traces: YSPEC (direct radial
# integration, full sphere, no attenuation, no gravity: test_01/yspec.in) for the same explosion,
# *velocity* of the moment-step response = displacement of the moment-rate delta response that
# AxiSEM's dirac_0 gives.  10 Hz, 18001 samples: smoothed with a 1.5 s Gaussian and decimated to
# 0.8 s (the comparison band ends at 0.05 Hz), Z and N only (E is zero for an explosion).
ys = read_mseed(os.path.join(ROOT, "test_01", "ref_data", "yspec.mseed"))
names = [l.split()[0] for l in open(os.path.join(ROOT, "test_01", "STATIONS")) if l.strip()]
sr, t0, d0 = ys[(names[0], "Z")]
assert abs(sr - 10.0) < 1e-9 and t0 == 0.0
tg = np.arange(-8.0, 8.0001, 0.1)
g = np.exp(-0.5 * (tg / 1.5) ** 2)
g /= g.sum()
dec = np.zeros((len(names), 2, len(d0[::8])), dtype=np.float32)
for i, n in enumerate(names):
    for c, comp in enumerate("NZ"):
        dec[i, c] = np.convolve(ys[(n, comp)][2], g, mode="same")[::8]
out["yspec_explosion_NZ"] = dec                                # (station, N/Z, sample)
out["yspec_dt"] = np.float64(0.8)
out["yspec_smoothing_sigma"] = np.float64(1.5)
print("yspec", dec.shape)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nightly_ref_seismograms.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path))
