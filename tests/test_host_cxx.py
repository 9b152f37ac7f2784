"""The native (C++) host of the seam `call time_loop` (axisem_b200/hostcxx/): module arrays
by their Fortran names -> C ABI -> chunked stepping -> receiver / wavefield buffers.

CPU: the host sources compiled against the oracle's implementation of the header must
reproduce the oracle driven through ctypes bit for bit (same library, different host), for
one rank and for two theta-slices in one process, and must fail the way the reference does.
GPU: the shipped executable (linked against libaxisem_b200.so) against the oracle within the
north_star tolerance (rel. L2 <= 1e-5, product build with FMA contraction)."""
import os
import subprocess

import numpy as np
import pytest

from axisem_b200.host.problem_bin import problem_records, save_problem_bin
from tests.util import make_problem, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_EXE = os.path.join(ROOT, "axisem_b200", "axisem_b200_solver")


def _run(exe, args):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)


def _oracle_exe():
    from oracle import oracle
    return oracle.build_host()


def _info(path):
    out = {}
    for line in open(path):
        k, v = line.split("=", 1)
        out[k.strip()] = v.strip()
    return out


def test_container_names_follow_the_fortran_modules():
    prob = make_problem("mtr", anel=True, niter=5, dump=True, strain_it=2)
    names = [n for n, _, _ in problem_records(prob)]
    assert len(names) == len(set(names))
    mods = {n.split("%")[0] for n in names}
    assert mods <= {"data_proc", "data_mesh", "data_spec", "data_matr", "data_pointwise", "data_source",
                    "data_time", "data_comm", "data_io", "attenuation"}
    for must in ("data_mesh%igloc_solid", "data_matr%M11s", "data_matr%M13s", "data_matr%inv_mass_rho",
                 "data_source%stf", "data_time%deltat", "data_mesh%recfile_el", "attenuation%y_j",
                 "data_matr%Y_cg4", "data_mesh%mapping_ijel_ikwf"):
        assert must in names, must
    assert "data_matr%M1phi" not in names          # quadrupole only (def_precomp_terms.f90:1216-1284)


def test_host_against_oracle_single_rank(tmp_path):
    from oracle import oracle
    n = 45
    prob = make_problem("mtr", anel=True, niter=n, dump=True, strain_it=10, seis_it=2)
    save_problem_bin(prob, str(tmp_path / "r0.axbp"))
    r = _run(_oracle_exe(), ["--out", str(tmp_path / "out"), "--dumpbuffer", "2", str(tmp_path / "r0.axbp")])
    assert r.returncode == 0, r.stderr
    assert "S T A R T I N G   T I M E   L O O P" in r.stdout
    info = _info(tmp_path / "out.info")
    O = oracle.make_loop(prob)
    O.run(n)
    assert int(info["iter"]) == n and int(info["nseismo"]) == O.nseismo and int(info["nstrain"]) == O.nstrain
    s = O.seismograms()
    got = np.fromfile(tmp_path / "out.rank0000.seis.f32", dtype=np.float32).reshape(s.shape)
    assert np.array_equal(got, s)
    sn = O.snapshots()
    got = np.fromfile(tmp_path / "out.rank0000.snap.f32", dtype=np.float32).reshape(sn.shape)
    assert np.array_equal(got, sn)          # chunked wavefield buffers reassembled in order


def test_host_writes_the_energy_table(tmp_path):
    from axisem_b200.host import SourceParams, build_problem
    from oracle import oracle
    from tests.util import small_spec
    n = 30
    prob = build_problem(small_spec(), SourceParams(src_type2="explosion", t_0=40.0), niter=n, energy=True)
    save_problem_bin(prob, str(tmp_path / "r0.axbp"))
    r = _run(_oracle_exe(), ["--quiet", "--out", str(tmp_path / "e"), str(tmp_path / "r0.axbp")])
    assert r.returncode == 0, r.stderr
    tab = np.loadtxt(tmp_path / "e.energy.txt")
    O = oracle.make_loop(prob)
    O.run(n)
    e = O.energy().astype(np.float64)
    assert tab.shape == (n + 1, 6)
    np.testing.assert_allclose(tab[:, 1:5], 2 * np.pi * e, rtol=2e-6)
    np.testing.assert_allclose(tab[:, 5], np.pi * e.sum(axis=1), rtol=2e-6)


def test_host_progress_lines_like_runtime_info(tmp_path):
    prob = make_problem("explosion", niter=230)
    save_problem_bin(prob, str(tmp_path / "r0.axbp"))
    r = _run(_oracle_exe(), ["--out", str(tmp_path / "out"), str(tmp_path / "r0.axbp")])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("  time step:")]
    # format 13 of runtime_info (time_evol_wave.F90:1017): every 100th step
    assert len(lines) == 2
    assert lines[0].startswith("  time step:   100; t=") and lines[0].rstrip().endswith("%)")


def test_host_two_slices_in_one_process(tmp_path):
    from axisem_b200.capi import connect_local, run_group
    from oracle import oracle
    n = 40
    probs = [make_problem("explosion", anel=True, niter=n, rank=r, nranks=2) for r in range(2)]
    files = []
    for r, p in enumerate(probs):
        files.append(str(tmp_path / f"p{r}.axbp"))
        save_problem_bin(p, files[-1])
    r = _run(_oracle_exe(), ["--quiet", "--out", str(tmp_path / "two")] + files)
    assert r.returncode == 0, r.stderr
    ol = [oracle.make_loop(p) for p in probs]
    olib = oracle.load()
    connect_local(olib, ol)
    run_group(olib, ol, n)
    for k, o in enumerate(ol):
        s = o.seismograms()
        got = np.fromfile(tmp_path / f"two.rank{k:04d}.seis.f32", dtype=np.float32).reshape(s.shape)
        assert np.array_equal(got, s)


def test_host_takes_the_mesh_from_the_meshers_database(tmp_path):
    """terms.axbp+meshdb.datNNNN: topology, spectral matrices and message lists come from the
    MESHER-format database through the native reader; the result is the same run."""
    from axisem_b200.capi import connect_local, run_group
    from axisem_b200.host.meshdb_io import write_meshdb
    from oracle import oracle
    n = 40
    probs = [make_problem("mtr", anel=True, niter=n, rank=r, nranks=2) for r in range(2)]
    args = []
    for r, p in enumerate(probs):
        save_problem_bin(p, str(tmp_path / f"t{r}.axbp"), without_mesh=True)
        write_meshdb(p.mesh, str(tmp_path / f"meshdb.dat{r:04d}"), dt=p.deltat)
        args.append(f"{tmp_path}/t{r}.axbp+{tmp_path}/meshdb.dat{r:04d}")
    r = _run(_oracle_exe(), ["--quiet", "--out", str(tmp_path / "m")] + args)
    assert r.returncode == 0, r.stderr
    ol = [oracle.make_loop(p) for p in probs]
    olib = oracle.load()
    connect_local(olib, ol)
    run_group(olib, ol, n)
    for k, o in enumerate(ol):
        s = o.seismograms()
        got = np.fromfile(tmp_path / f"m.rank{k:04d}.seis.f32", dtype=np.float32).reshape(s.shape)
        assert np.array_equal(got, s)
    # without the database the stripped container is not enough, and the host says what is missing
    r = _run(_oracle_exe(), ["--quiet", "--out", str(tmp_path / "x"), str(tmp_path / "t0.axbp")])
    assert r.returncode == 1 and "data_" in r.stderr


def test_host_error_behaviour(tmp_path):
    exe = _oracle_exe()
    # a file that is not a container: message + non-zero exit, like the reference's stop
    bad = tmp_path / "bad.axbp"
    bad.write_bytes(b"not a problem")
    r = _run(exe, ["--out", str(tmp_path / "o"), str(bad)])
    assert r.returncode == 1 and "ERROR" in r.stderr
    # a missing module variable is named
    prob = make_problem("explosion", niter=5)
    import struct
    recs = [x for x in problem_records(prob) if x[0] != "data_matr%inv_mass_rho"]
    from axisem_b200.host import problem_bin
    out = [b"AXBPROB1", struct.pack("<I", len(recs))]
    for name, a, t in recs:
        problem_bin._rec(out, name, a, t)
    (tmp_path / "miss.axbp").write_bytes(b"".join(out))
    r = _run(exe, ["--out", str(tmp_path / "o"), str(tmp_path / "miss.axbp")])
    assert r.returncode == 1 and "inv_mass_rho" in r.stderr
    # more steps than niter
    save_problem_bin(prob, str(tmp_path / "ok.axbp"))
    r = _run(exe, ["--steps", "6", "--out", str(tmp_path / "o"), str(tmp_path / "ok.axbp")])
    assert r.returncode == 1 and "niter" in r.stderr


@pytest.mark.gpu
def test_product_host_on_gpu(tmp_path):
    from oracle import oracle
    assert os.path.exists(PRODUCT_EXE), "axisem_b200_solver missing: run __graft_entry__.build()"
    n = 60
    prob = make_problem("mtr", anel=True, niter=n, dump=True, strain_it=15)
    save_problem_bin(prob, str(tmp_path / "r0.axbp"))
    r = _run(PRODUCT_EXE, ["--out", str(tmp_path / "out"), "--dumpbuffer", "3", str(tmp_path / "r0.axbp")])
    assert r.returncode == 0, r.stderr
    info = _info(tmp_path / "out.info")
    assert int(info["gpu_launches"]) > 0 and int(info["iter"]) == n
    O = oracle.make_loop(prob)
    O.run(n)
    s = O.seismograms()
    got = np.fromfile(tmp_path / "out.rank0000.seis.f32", dtype=np.float32).reshape(s.shape)
    assert rel_l2(got, s) <= 1e-5
    sn = O.snapshots()
    got = np.fromfile(tmp_path / "out.rank0000.snap.f32", dtype=np.float32).reshape(sn.shape)
    assert rel_l2(got, sn) <= 1e-5


@pytest.mark.gpu
def test_product_host_stops_on_blowup(tmp_path):
    """runtime_info's guard (time_evol_wave.F90:1042-1054): |disp(1,1,:,:)| > 10 |magnitude| ends the
    run with the reference's message and a non-zero exit."""
    prob = make_problem("explosion", niter=300)
    prob.source.magnitude = 1e-30           # the stf keeps its size: the threshold is exceeded at once
    save_problem_bin(prob, str(tmp_path / "blow.axbp"))
    r = _run(PRODUCT_EXE, ["--quiet", "--out", str(tmp_path / "o"), str(tmp_path / "blow.axbp")])
    assert r.returncode == 1 and "BLEW UP" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_product_host_two_slices_on_gpu(tmp_path):
    """Both theta-slices driven by one process (one handle per slice; on a multi-GPU box pass
    --devices N and the peers are wired over NVLink)."""
    import torch
    from axisem_b200.capi import connect_local, run_group
    from oracle import oracle
    n = 30
    probs = [make_problem("mtr", anel=True, niter=n, rank=r, nranks=2) for r in range(2)]
    files = []
    for r, p in enumerate(probs):
        files.append(str(tmp_path / f"p{r}.axbp"))
        save_problem_bin(p, files[-1])
    ndev = min(2, torch.cuda.device_count())
    r = _run(PRODUCT_EXE, ["--quiet", "--devices", str(ndev), "--out", str(tmp_path / "two")] + files)
    assert r.returncode == 0, r.stderr
    ol = [oracle.make_loop(p) for p in probs]
    olib = oracle.load()
    connect_local(olib, ol)
    run_group(olib, ol, n)
    for k, o in enumerate(ol):
        s = o.seismograms()
        got = np.fromfile(tmp_path / f"two.rank{k:04d}.seis.f32", dtype=np.float32).reshape(s.shape)
        assert rel_l2(got, s) <= 1e-5
