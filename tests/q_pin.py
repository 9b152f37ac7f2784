"""An oracle-independent pin of the anelastic loop: the attenuation a P wave actually suffers in a
homogeneous anelastic sphere, against the quality factor of the standard-linear-solid set.

The reference ships no anelastic traces (TESTING/nightly has elastic cases only), so this is
the check that ties the memory-variable machinery — a_j, the delta moduli, the unrelaxed moduli
(attenuation.f90:1009-1063), the update recursion (:136-200) and the anelastic stiffness terms —
to an analytic statement: the complex modulus of the SLS set (Emmerich & Korn 1987, the formula
behind the reference's own `q_linear_solid`, attenuation.f90:1095-1135).

Set-up: one solid layer (hollow below 1500 km), explosion at 1200 km depth, two surface stations
at 35 and 80 degrees; the same run with and without attenuation.  With X_k the spectrum of the
windowed direct P wave at station k,

    H(w) = [X_2 / X_1]_anelastic / [X_2 / X_1]_elastic = exp(-i w (s(w) - 1/v_p) (L_2 - L_1))

(source spectrum, radiation, geometrical spreading and free-surface factors cancel), so
-ln|H| / (w dL) is the imaginary part of the complex slowness s = sqrt(rho / M_P(w)), M_P =
kappa(w) + 4/3 mu(w), i.e. 1 / (2 c Q_P), and -arg H / (w dL) its real part minus 1/v_p
(dispersion).

Two theories are compared with the measurement:
  * `continuous`: M(w) = M_u - dM sum_j a_j w_j / (i w + w_j)  — exactly q_linear_solid(exact);
  * `discrete`: the same with the transfer function of the loop's recursion
    R(n+1) = e R(n) + c_t s(n+1) + c_tm1 s(n) and the one-step lag with which the reference
    applies it (the anelastic stiffness of step n+1 uses R(n): time_evol_wave.F90:395-455).  The lag
    adds attenuation in proportion to w dt — 29 % at dt = 0.48 s, 17 % at half of that — so the
    continuous value is approached linearly in dt.
"""
from __future__ import annotations

import numpy as np

from axisem_b200.host import AttenuationModel, SourceParams, build_problem, homogeneous_layers
from axisem_b200.host.mesh import MeshSpec
from axisem_b200.host.precomp import fast_correct

R, RMIN = 6371e3, 1500e3
VP, VS, RHO = 8000.0, 4500.0, 3000.0
T0, DEPTH = 40.0, 1200e3
COLAT = np.array([35.0, 80.0])
NT, NR = 128, 32
T_END = 1100.0


def problems(qmu, qka, anel, cg, nranks, courant=0.6):
    spec = MeshSpec(ntheta=NT, layers=homogeneous_layers(r_min_km=RMIN / 1e3, rho=RHO, vp=VP, vs=VS,
                                                        qmu=qmu, qkappa=qka), nrad=[NR])
    sp = SourceParams(src_type2="explosion", depth=DEPTH, t_0=T0)
    att = AttenuationModel(coarse_grained=cg) if anel else None
    dt = build_problem(spec, sp, anel=anel, att=att, niter=4, rec_colat_deg=COLAT, courant=courant).deltat
    niter = int(T_END / dt)
    return [build_problem(spec, sp, anel=anel, att=att, niter=niter, rec_colat_deg=COLAT, rank=r,
                          nranks=nranks, courant=courant) for r in range(nranks)], niter, dt


def gather(probs, loops, niter):
    s = np.zeros((niter + 1, COLAT.size, 3))
    for p, L in zip(probs, loops):
        if p.num_rec:
            s[:, p.rec_index, :] = L.seismograms()
    return s


def complex_moduli(att: AttenuationModel, qmu, qka, w, dt=None):
    """kappa(w), mu(w) of the SLS set as the loop realises it (dt None: continuous limit)."""
    w_j, y_j = np.asarray(att.w_j, float), np.asarray(att.y_j, float)
    w_0 = att.w_0 * 2 * np.pi
    w_1 = np.sqrt(att.f_min * att.f_max) * 2 * np.pi
    mu = RHO * VS ** 2
    ka = RHO * VP ** 2 - 4.0 / 3.0 * mu

    def modulus(M, Q):
        yp = fast_correct(y_j / Q) if att.do_corr_lowq else y_j / Q
        fac = np.sum(yp * w_j ** 2 / (w_1 ** 2 + w_j ** 2)) / yp.sum()
        M_w1 = M * (1 + 2.0 / (np.pi * Q) * np.log(w_1 / w_0))
        dM = M_w1 / (1.0 / yp.sum() + 1 - fac)
        M_u = M_w1 + dM * fac
        a = yp / yp.sum()
        if dt is None:
            G = w_j[None, :] / (1j * w[:, None] + w_j[None, :])
        else:
            e = np.exp(-w_j * dt)
            c_tm1 = (1 - e) / (w_j * dt) - e
            c_t = (e - 1) / (w_j * dt) + 1
            z = np.exp(1j * w * dt)[:, None]
            G = (c_t[None, :] + c_tm1[None, :] / z) / (z - e[None, :])
        return M_u - dM * np.sum(a[None, :] * G, axis=1)

    return modulus(ka, qka), modulus(mu, qmu)


def q_linear_solid(y_j, w_j, w):
    """attenuation.f90:1095-1135 with exact = .true."""
    num = 1 + np.sum(y_j[None, :] * w[:, None] ** 2 / (w[:, None] ** 2 + w_j[None, :] ** 2), axis=1)
    den = np.sum(y_j[None, :] * w[:, None] * w_j[None, :] / (w[:, None] ** 2 + w_j[None, :] ** 2), axis=1)
    return num / den


def measure(s_el, s_an, dt):
    """Imaginary part and excess real part of the P slowness from the two-station spectral ratio.
    Returns (frequencies, weights = normalised source spectrum, Im s (positive), Re s - 1/vp)."""
    n = s_el.shape[0]
    t = np.arange(n) * dt
    shift = np.ceil(1.5 * T0 / dt) * dt
    rs = R - DEPTH
    L = np.sqrt(R ** 2 + rs ** 2 - 2 * R * rs * np.cos(np.deg2rad(COLAT)))
    nf = 1 << 16
    fr = np.fft.rfftfreq(nf, dt)

    def radial(s, k):
        th = np.deg2rad(COLAT[k])
        return s[:, k, 0] * np.sin(th) + s[:, k, 2] * np.cos(th)

    def spectrum(x, tc):
        m0 = (t > tc - 2 * T0) & (t < tc + 3 * T0)             # centre on the arrival's energy:
        tcen = np.sum(t[m0] * x[m0] ** 2) / np.sum(x[m0] ** 2)  # the anelastic pulse comes later
        w0, w1 = tcen - 2.2 * T0, tcen + 2.2 * T0
        m = (t > w0) & (t < w1)
        tt = (t[m] - w0) / (w1 - w0)
        tap = np.ones(m.sum())
        e = 0.15
        tap[tt < e] = 0.5 * (1 - np.cos(np.pi * tt[tt < e] / e))
        tap[tt > 1 - e] = 0.5 * (1 - np.cos(np.pi * (1 - tt[tt > 1 - e]) / e))
        win = np.zeros(n)
        win[m] = tap
        return np.fft.rfft(x * win, nf)

    X = {nm: [spectrum(radial(s, k), L[k] / VP + shift) for k in range(2)] for nm, s in (("el", s_el), ("an", s_an))}
    H = (X["an"][1] / X["an"][0]) / (X["el"][1] / X["el"][0])
    dL = L[1] - L[0]
    band = (fr > 0.008) & (fr < 0.05)
    w = 2 * np.pi * fr[band]
    A = np.abs(X["el"][0][band])
    return fr[band], A / A.max(), -np.log(np.abs(H[band])) / (w * dL), -np.unwrap(np.angle(H[band])) / (w * dL)


def compare(s_el, s_an, dt, qmu, qka, discrete=True, att=None):
    """Energy-weighted mean over the band carrying at least half of the peak source spectrum of
    measured / predicted Im s (attenuation) and Re s - 1/vp (dispersion)."""
    att = att or AttenuationModel()
    fr, A, im_meas, re_meas = measure(s_el, s_an, dt)
    ka, mu = complex_moduli(att, qmu, qka, 2 * np.pi * fr, dt if discrete else None)
    s_th = np.sqrt(RHO / (ka + 4.0 / 3.0 * mu))
    sel = A > 0.5
    wgt = A[sel] / A[sel].sum()
    return (float(np.sum(wgt * im_meas[sel] / -np.imag(s_th[sel]))),
            float(np.sum(wgt * re_meas[sel] / (np.real(s_th[sel]) - 1.0 / VP))))
