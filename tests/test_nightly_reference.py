"""End-to-end pin against the reference's own golden seismograms:
TESTING/nightly/test_0{1,2,3}/ref_data/axisem.mseed — explosion (monopole), mtr (dipole), mtp
(quadrupole) in elastic prem_ani, the traces the reference's nightly regression compares every
build against — committed as tests/golden/nightly_ref_seismograms.npz.

See tests/nightly_compare.py for what is compared and why the agreement is at waveform level
(correlation, amplitude ratio) rather than sample level: the meshes differ, and source and
receivers sit on the nearest GLL points of either mesh.

CPU: the oracle (8 theta-slices in threads) on a coarse mesh, dipole source, all of E, N, Z.
GPU: the CUDA library on a finer mesh, all three source types."""
import os

import numpy as np
import pytest

from axisem_b200.host import SourceParams, build_problem, prem_mesh_spec
from tests.nightly_compare import compare, stations, to_enz

T_0 = 70.0          # Gaussian source period [s]: keeps the comparison inside the band both meshes resolve


# The reference puts the source on the GLL point nearest to the requested 100 km
# (find_srcloc, source.f90:454-476).  Its 50 s mesh (1.5 elements per wavelength) has one element
# between the 80 km and 220 km discontinuities, whose GLL radii are 6151 + 140 (0, 0.1727, 0.5,
# 0.8273, 1) km: the point nearest to 100 km depth lies at 104.2 km.  (Inferred from the mesher's
# rules, not read from a mesh; with the source at 100 km instead, Rayleigh-wave amplitudes at
# near stations come out 10 % high.)  The meshes here get a GLL point within 1 km of that depth.
SOURCE_DEPTH = 104.2e3


def _setup(src, ntheta, nr, nranks, scheme="newmark2", lvz_elements=1):
    from axisem_b200.host.mesh import MeshSpec
    names, lat, lon = stations()
    colat = 90.0 - lat
    base = prem_mesh_spec(ntheta=ntheta, nr_target=nr, anisotropic=True, r_min_km=800.0)
    nrad = list(base.nrad)
    nrad[[L.name for L in base.layers].index("LVZ")] = lvz_elements      # 1 or 3: GLL point at 104.2 / 103.3 km
    spec = MeshSpec(ntheta=ntheta, layers=base.layers, nrad=nrad)
    sp = SourceParams(src_type2=src, depth=SOURCE_DEPTH, magnitude=1e20, t_0=T_0)
    dt = build_problem(spec, sp, niter=4, rec_colat_deg=colat, time_scheme=scheme).deltat
    shift = np.ceil(1.5 * T_0 / dt) * dt
    niter = int((1800.0 + shift) / dt) + 1
    seis_it = max(1, int(0.8 / dt))
    probs = [build_problem(spec, sp, niter=niter, rec_colat_deg=colat, seis_it=seis_it, rank=r, nranks=nranks,
                           time_scheme=scheme) for r in range(nranks)]
    return probs, niter, seis_it, dt, shift, np.deg2rad(colat), np.deg2rad(lon)


def _score(src, loops, probs, niter, seis_it, dt, shift, colat, lon, against="axisem"):
    ns = max(L.nseismo for L in loops)
    s = np.zeros((ns, colat.size, 3))
    for p, L in zip(probs, loops):
        if p.num_rec:
            s[:, p.rec_index, :] = L.seismograms()
    t = np.arange(ns) * seis_it * dt - shift                 # origin time = centre of the Gaussian
    r = compare(src, to_enz(src, s, colat, lon), t, T_0, against=against)
    big = r[:, :, 3] > 0.002 * r[:, :, 3].max()              # traces that carry signal (explosion: E is zero)
    return r[:, :, 0][big], r[:, :, 2][big], r


@pytest.fixture(scope="module")
def native_oracle_run(tmp_path_factory):
    """The dipole case through the native pipeline on the CPU: AXBPROB1 containers of 8
    theta-slices -> the C++ host (axisem_b200/hostcxx, compiled against the oracle's
    implementation of the header) -> RUN.rankNNNN.seis.f32."""
    import subprocess
    from axisem_b200.host.problem_bin import save_problem_bin
    from oracle import oracle
    tmp = tmp_path_factory.mktemp("nightly")
    nranks = min(8, os.cpu_count() or 1)
    probs, niter, *rest = _setup("mtr", 128, 40, nranks)
    files = []
    for r, p in enumerate(probs):
        files.append(str(tmp / f"r{r}.axbp"))
        save_problem_bin(p, files[-1])
    run = subprocess.run([oracle.build_host(), "--quiet", "--out", str(tmp / "run")] + files,
                         capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr
    return tmp, probs, niter, rest


def _native_seismograms(tmp, probs, nsta):
    s = None
    for r, p in enumerate(probs):
        if not p.num_rec:
            continue
        raw = np.fromfile(tmp / f"run.rank{r:04d}.seis.f32", dtype=np.float32)
        raw = raw.reshape(-1, p.num_rec, 3)                     # recdumpvar(3, num_rec, nseismo)
        if s is None:
            s = np.zeros((raw.shape[0], nsta, 3))
        s[:, p.rec_index, :] = raw
    return s


def test_oracle_reproduces_the_references_dipole_seismograms(native_oracle_run):
    tmp, probs, niter, (seis_it, dt, shift, colat, lon) = native_oracle_run
    s = _native_seismograms(tmp, probs, colat.size)
    t = np.arange(s.shape[0]) * seis_it * dt - shift
    r = compare("mtr", to_enz("mtr", s, colat, lon), t, T_0)
    big = r[:, :, 3] > 0.002 * r[:, :, 3].max()
    cc, amp = r[:, :, 0][big], r[:, :, 2][big]
    assert cc.size >= 40                                      # of 20 stations x (E, N, Z)
    assert cc.min() > 0.85 and np.median(cc) > 0.99, (cc.min(), np.median(cc))       # measured 0.878 / 0.9958
    assert amp.min() > 0.93 and amp.max() < 1.08, (amp.min(), amp.max())              # measured 0.959 .. 1.049
    assert abs(np.median(amp) - 1.0) < 0.02


def test_native_postprocessing_of_the_same_run(native_oracle_run):
    """axisem_b200_postproc (radiation factors, rotation into E, N, Z: hostcxx/postprocess.cpp)
    on the files the C++ host wrote: equal to the Python restatement, and therefore as close to
    the reference's traces."""
    import subprocess
    tmp, probs, niter, (seis_it, dt, shift, colat, lon) = native_oracle_run
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "axisem_b200", "axisem_b200_postproc")
    if not os.path.exists(exe):
        subprocess.check_call(["bash", os.path.join(root, "axisem_b200", "hostcxx", "build.sh")])
    s = _native_seismograms(tmp, probs, colat.size)
    want = to_enz("mtr", s, colat, lon)[:, [1, 0, 2], :]      # (station, N/E/Z, sample): the reference's order
    got = np.zeros_like(want)
    for r, p in enumerate(probs):
        if not p.num_rec:
            continue
        st = tmp / f"st{r}.txt"
        st.write_text("".join(f"{np.rad2deg(colat[k]):.10f} {np.rad2deg(lon[k]):.10f}\n" for k in p.rec_index))
        out = tmp / f"enz{r}.f32"
        run = subprocess.run([exe, "--src", "mtr", "--sys", "enz", "--stations", str(st), "--seis",
                              str(tmp / f"run.rank{r:04d}.seis.f32"), "--out", str(out)], capture_output=True, text=True)
        assert run.returncode == 0, run.stderr
        got[p.rec_index] = np.fromfile(out, dtype=np.float32).reshape(p.num_rec, 3, -1)
    # the reference recovers the receiver's longitude through acos((x + 1e-11) / (|xy| + 1e-11))
    # (post_processing.F90:214-224), good to ~5e-6 rad near phi = 0: that bounds the agreement
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(got - want) <= 1e-5 * scale + 1e-30)
    # Gaussian convolution of the tool against numpy's
    k = int(probs[0].rec_index[0]) if probs[0].num_rec else 0
    src_rank = next(r for r, p in enumerate(probs) if p.num_rec)
    p = probs[src_rank]
    out = tmp / "conv.f32"
    sdt = seis_it * dt
    run = subprocess.run([exe, "--src", "mtr", "--sys", "cyl", "--conv", "60.0", "3.5", f"{sdt:.12f}", "--stations",
                          str(tmp / f"st{src_rank}.txt"), "--seis", str(tmp / f"run.rank{src_rank:04d}.seis.f32"),
                          "--out", str(out)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    conv = np.fromfile(out, dtype=np.float32).reshape(p.num_rec, 3, -1)
    a = 3.5 / 60.0
    half = int(np.ceil(4 * 60.0 / sdt))
    g = a / np.sqrt(np.pi) * np.exp(-(a * np.arange(-half, half + 1) * sdt) ** 2) * sdt
    st0 = int(p.rec_index[0])
    raw = np.fromfile(tmp / f"run.rank{src_rank:04d}.seis.f32", dtype=np.float32).reshape(-1, p.num_rec, 3)
    z = raw[:, 0, 2].astype(np.float64) * np.cos(lon[st0])    # mtr: the z factor is cos(phi)
    ref = np.convolve(z, g)[half:half + z.size]
    assert np.allclose(conv[0, 2], ref, rtol=0, atol=3e-6 * np.abs(ref).max())


@pytest.mark.parametrize("src,bar", [("explosion", (0.88, 0.97, 0.80, 1.12)), ("mtp", (0.88, 0.975, 0.93, 1.13))])
def test_oracle_reproduces_the_references_monopole_and_quadrupole_seismograms(src, bar):
    """The other two source orders on the coarse mesh (oracle through ctypes, 8 theta-slices in
    threads).  Measured: explosion 0.921 / 0.981, amplitude 0.85 .. 1.07 (median 1.003);
    mtp 0.918 / 0.986, 0.98 .. 1.09 (median 1.002).  The fine-mesh numbers are in the GPU tests."""
    from axisem_b200.capi import TimeLoop, connect_local, run_group
    from oracle import oracle
    nranks = min(8, os.cpu_count() or 1)
    probs, niter, *rest = _setup(src, 128, 40, nranks)
    lib = oracle.load_fast()
    loops = [TimeLoop(lib, p) for p in probs]
    if nranks > 1:
        connect_local(lib, loops)
        run_group(lib, loops, niter)
    else:
        loops[0].run(niter)
    cc, amp, _ = _score(src, loops, probs, niter, *rest)
    cmin, cmed, alo, ahi = bar
    assert cc.size >= 35
    assert cc.min() > cmin and np.median(cc) > cmed, (cc.min(), np.median(cc))
    assert amp.min() > alo and amp.max() < ahi, (amp.min(), amp.max())
    assert abs(np.median(amp) - 1.0) < 0.02


# measured on the 224 x 60 mesh (oracle and CUDA library alike): correlation min / median, amplitude range
#   explosion 0.921 / 0.9985, 0.84 .. 1.07 (median 1.006)   [the low ones are small core phases at > 130 degrees]
#   mtr       0.991 / 0.9997, 0.97 .. 1.04 (median 1.005)
#   mtp       0.996 / 0.9998, 0.99 .. 1.02 (median 1.003)
BAR = {"explosion": (0.90, 0.997, 0.80, 1.10), "mtr": (0.98, 0.999, 0.95, 1.06), "mtp": (0.98, 0.999, 0.95, 1.06)}


@pytest.mark.gpu
@pytest.mark.parametrize("src", ["explosion", "mtr", "mtp"])
def test_cuda_reproduces_the_references_seismograms(src):
    from axisem_b200 import solver
    probs, niter, *rest = _setup(src, 224, 60, 1, lvz_elements=3)
    loop = solver.time_loop(probs[0])
    loop.run(niter)
    cc, amp, r = _score(src, [loop], probs, niter, *rest)
    print(f"{src}: {cc.size} traces, correlation min {cc.min():.4f} median {np.median(cc):.4f}, "
          f"amplitude ratio {amp.min():.3f} .. {amp.max():.3f}, launches {loop.gpu_launches}")
    assert loop.gpu_launches > 0
    cmin, cmed, alo, ahi = BAR[src]
    assert cc.size >= 35
    assert cc.min() > cmin and np.median(cc) > cmed, (cc.min(), np.median(cc))
    assert amp.min() > alo and amp.max() < ahi, (amp.min(), amp.max())
    assert abs(np.median(amp) - 1.0) < 0.02, np.median(amp)
    if src == "explosion":
        # the independent YSPEC solution (full sphere, no attenuation, no gravity) that the
        # reference ships next to its own traces: test_01/ref_data/yspec.mseed
        cy, ay, _ = _score(src, [loop], probs, niter, *rest, against="yspec")
        print(f"  against yspec: {cy.size} traces, correlation min {cy.min():.4f} median {np.median(cy):.4f}, "
              f"amplitude ratio {ay.min():.3f} .. {ay.max():.3f}")
        assert cy.size >= 35 and cy.min() > 0.90 and np.median(cy) > 0.997, (cy.min(), np.median(cy))   # 0.920 / 0.9984
        assert ay.min() > 0.80 and ay.max() < 1.12 and abs(np.median(ay) - 1.0) < 0.02, (ay.min(), ay.max())


@pytest.mark.gpu
@pytest.mark.timeout(1800)
@pytest.mark.parametrize("src", ["explosion", "mtr", "mtp"])
def test_cuda_equals_the_oracle_over_the_whole_nightly_run(src):
    """The sentence of DESIGN.md section 2 ("identical for the oracle and the CUDA library") as an
    assertion: the same 224 x 60 mesh, all 18 293 steps of the 1800 s run.  The oracle runs as 8
    theta-slices in threads; the strict build runs the same 8 slices (loop-back halo on one GPU)
    and must be bit-identical in every seismogram sample; the product build (one slice, FMA
    contraction, lean Newmark, graph replay) must stay within 1e-5 relative L2."""
    from axisem_b200 import solver
    from axisem_b200.capi import TimeLoop, connect_local, run_group
    from oracle import oracle
    from tests.util import rel_l2
    nranks = 8
    probs, niter, *rest = _setup(src, 224, 60, nranks, lvz_elements=3)
    colat = rest[3]
    olib = oracle.load()
    ol = [TimeLoop(olib, p) for p in probs]
    connect_local(olib, ol)
    run_group(olib, ol, niter)

    def gather(loops, ps):
        ns = max(L.nseismo for L in loops)
        s = np.zeros((ns, colat.size, 3), dtype=np.float32)
        for p, L in zip(ps, loops):
            if p.num_rec:
                s[:, p.rec_index, :] = L.seismograms()
        return s

    ref = gather(ol, probs)
    del ol
    assert np.abs(ref).max() > 0
    lib, gl = solver.time_loop_group(probs, strict=True)
    run_group(lib, gl, niter)
    for g in gl:
        g.synchronize()
    strict = gather(gl, probs)
    del gl
    assert np.array_equal(strict, ref), f"strict build, 8 slices: rel l2 {rel_l2(strict, ref):.3e}"
    one, niter1, *_ = _setup(src, 224, 60, 1, lvz_elements=3)
    assert niter1 == niter
    P = solver.time_loop(one[0])
    P.run(niter)
    prod = gather([P], one)
    err = rel_l2(prod, ref)
    print(f"{src}: {niter} steps, product (1 slice) vs oracle (8 slices): rel l2 {err:.3e}; launches {P.gpu_launches}")
    assert err <= 1e-5, err


@pytest.mark.gpu
def test_cuda_symplectic_scheme_reproduces_the_references_dipole_seismograms():
    """The 4th-order symplectic loop (time step 1.5 x Newmark's, point-wise source time function
    of compute_stf_t) against the same golden traces; the oracle gives the same numbers."""
    from axisem_b200 import solver
    probs, niter, *rest = _setup("mtr", 128, 40, 1, scheme="symplec4")
    loop = solver.time_loop(probs[0])
    loop.run(niter)
    cc, amp, _ = _score("mtr", [loop], probs, niter, *rest)
    print(f"symplec4 mtr: {cc.size} traces, correlation min {cc.min():.4f} median {np.median(cc):.4f}, "
          f"amplitude ratio {amp.min():.3f} .. {amp.max():.3f}")
    assert cc.size >= 40
    assert cc.min() > 0.85 and np.median(cc) > 0.99, (cc.min(), np.median(cc))       # measured 0.878 / 0.9957
    assert amp.min() > 0.93 and amp.max() < 1.08, (amp.min(), amp.max())


@pytest.mark.gpu
def test_the_two_attenuation_formulations_agree_with_each_other():
    """No reference trace exists for the anelastic loop, so its two independent formulations
    check each other on the nightly dipole set-up: coarse-grained memory variables at 4 points
    (attenuation.f90:81-202, stiffness_*:glob_anel_stiffness_*_cg4) against memory variables at
    all 25 points (:210-334, glob_anel_stiffness_*_4) — different arrays, different kernels
    (fused in S_A / k_anel_full).  Measured with the oracle on this mesh: relative L2 difference
    of the seismograms 0.040 between the two, 0.7 between either and the elastic run."""
    from axisem_b200 import solver
    from axisem_b200.host import AttenuationModel
    from tests.util import rel_l2
    t_0 = 100.0
    colat = np.array([20.0, 50.0, 80.0, 110.0, 140.0, 170.0])
    spec = prem_mesh_spec(ntheta=96, nr_target=32, anisotropic=True, r_min_km=1000.0)
    sp = SourceParams(src_type2="mtr", depth=100e3, magnitude=1e20, t_0=t_0)
    out = {}
    for key, anel, cg in (("elastic", False, True), ("cg4", True, True), ("full", True, False)):
        att = AttenuationModel(coarse_grained=cg) if anel else None
        dt = build_problem(spec, sp, niter=4, rec_colat_deg=colat, anel=anel, att=att).deltat
        shift = np.ceil(1.5 * t_0 / dt) * dt
        niter = int((1800.0 + shift) / dt) + 1
        prob = build_problem(spec, sp, niter=niter, rec_colat_deg=colat, seis_it=max(1, int(2.0 / dt)),
                             anel=anel, att=att)
        loop = solver.time_loop(prob)
        loop.run(niter)
        out[key] = loop.seismograms().astype(np.float64)
        assert np.isfinite(out[key]).all() and np.abs(out[key]).max() > 0
    d_formulations = rel_l2(out["cg4"], out["full"])
    d_physics = rel_l2(out["cg4"], out["elastic"])
    print(f"cg4 vs full memory variables: {d_formulations:.4f}; anelastic vs elastic: {d_physics:.3f}")
    assert d_formulations < 0.08, d_formulations
    assert d_physics > 0.4, d_physics
    peak = np.abs(out["cg4"]).max(axis=(0, 2)) / np.abs(out["elastic"]).max(axis=(0, 2))
    assert np.all(peak < 1.0) and np.all(peak > 0.4), peak          # attenuation lowers every station's peak
