#!/usr/bin/env python
"""bench.py — GLL-point updates/s of the device-resident AxiSEM time loop on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            (N=1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N>1)
    python bench.py --impl reference ...     reference-equivalent CPU path on the host cores

Workload (BASELINE.json configs[4], the only configuration whose per-GPU footprint is far
above the 126 MB L2): a synthetic PREM-type mesh of ~4.0e6 spectral elements = 1.0e8
element-local GLL points (npol=4), dipole (mtr) moment-tensor source, coarse-grained
attenuation with 5 SLS, Newmark; theta-sliced over the N GPUs of one box (strong scaling:
the mesh is fixed, every rank owns ntheta/N columns).  One "step" = one full Newmark time
step of the coupled solid/fluid system over the whole mesh.

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from axisem_b200.host import AttenuationModel, SourceParams, build_problem, prem_mesh_spec  # noqa: E402

# algorithmic bytes per element-local GLL point and time step
B_SOLID = {"monopole": 112.0, "dipole": 172.0, "quadpole": 180.0}
B_FLUID = {"monopole": 44.0, "dipole": 48.0, "quadpole": 48.0}
B_ANEL = {"monopole": 40.0, "dipole": 55.0, "quadpole": 55.0}
B_ANEL_FULL = {"monopole": 248.0, "dipole": 344.0, "quadpole": 344.0}   # COARSE_GRAINED false
# the share of those bytes that belongs to the solid element kernel S_A:
# disp r+w (2 nc) + all M planes (ncoef - inv_mass) [+ the whole anelastic part]
NC = {"monopole": 2, "dipole": 3, "quadpole": 3}
NCOEF = {"monopole": 15, "dipole": 24, "quadpole": 26}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=5,
                    help="timed regions of --steps steps each; the line reports their median and spread")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ntheta", type=int, default=3584)
    ap.add_argument("--nr", type=int, default=1116)
    ap.add_argument("--source", default="mtr")
    ap.add_argument("--no-anel", action="store_true")
    ap.add_argument("--full-memvars", action="store_true",
                    help="COARSE_GRAINED false: memory variables at all 25 points (not the headline config)")
    ap.add_argument("--cpu-sample-cols", type=int, default=0,
                    help="theta columns of the CPU-baseline mesh (same radial structure); 0 = the named "
                         "mesh itself when the host has the memory for it, else half of it")
    ap.add_argument("--cpu-steps", type=int, default=0, help="0 = size for ~15 s")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true",
                    help="skip the correctness check of the N-rank run against the committed 1-rank golden")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(args, src, anel, cores=None):
    """The oracle (line-by-line CPU restatement of the reference loop) compiled -O3 with FMA
    contraction (oracle/libaxisem_oracle_fast.so: the way a production CPU build would be; the
    parity checks use the strict build), one thread per theta-slice = one 'MPI rank' per host
    core, on a bounded sample of the workload: the same radial structure and physics, fewer
    theta columns."""
    from oracle import oracle
    from axisem_b200.capi import connect_local, run_group
    cores = cores or (os.cpu_count() or 1)
    want = args.cpu_sample_cols
    if want <= 0:
        # the oracle holds about 7 KB per element next to the host arrays it was given (same again)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 0
        need_full = 16e3 * args.ntheta * args.nr
        want = args.ntheta if avail > 1.5 * need_full else max(args.ntheta // 2, 2 * cores)
    ncols = max(want // cores, 2) * cores
    if ncols % 2:                                      # the mesh generator wants an even count
        ncols += cores
    spec = prem_mesh_spec(ntheta=ncols, nr_target=args.nr)
    lib = oracle.load_fast()
    nsteps = args.cpu_steps or 3
    att = AttenuationModel(coarse_grained=False) if (anel and args.full_memvars) else None
    probs = [build_problem(spec, SourceParams(src_type2=src), anel=anel, att=att, niter=400, rank=r,
                           nranks=cores, rec_colat_deg=[]) for r in range(cores)]
    from axisem_b200.capi import TimeLoop
    loops = [TimeLoop(lib, p) for p in probs]
    rng = np.random.default_rng(1234)
    for L in loops:
        for f in ("disp", "velo"):
            L.set(f, (rng.standard_normal(L._field_shape(f)) * 1e-6).astype(np.float32))
    if cores > 1:
        connect_local(lib, loops)
    run = (lambda n: run_group(lib, loops, n)) if cores > 1 else (lambda n: loops[0].run(n))
    run(1)                                           # warm-up
    t = time.perf_counter()
    run(nsteps)
    dt = time.perf_counter() - t
    if not args.cpu_steps:                           # size the sample for ~15 s of CPU work
        more = int(min(max(12.0 / (dt / nsteps) - nsteps, 0), 390 - nsteps))
        if more > 0:
            t2 = time.perf_counter()
            run(more)
            dt += time.perf_counter() - t2
            nsteps += more
    pts = 25 * spec.nelem
    return {"value": pts * nsteps / dt, "unit": "GLL-point updates/s", "cores": cores,
            "kind": "port", "same_mesh": bool(ncols == args.ntheta),
            "sample": f"{spec.nelem} elements ({ncols} theta columns x {spec.nr} radial, same "
                      f"layering/physics), {nsteps} steps, {dt:.1f} s",
            "ms_per_step": dt / nsteps * 1e3, "points": pts, "steps": nsteps}


def main():
    args = parse()
    src = args.source
    anel = not args.no_anel
    sp = SourceParams(src_type2=src)
    pole = sp.src_type1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.gpus
    workload = (f"PREM-type synthetic mesh {args.ntheta}x{args.nr} (theta x r), {pole} ({src}) "
                f"source, {('full (25-point) attenuation 5 SLS' if args.full_memvars else 'cg4 attenuation 5 SLS') if anel else 'elastic'}, newmark2")

    if args.impl == "reference":
        if rank != 0:
            return
        spec = prem_mesh_spec(ntheta=args.ntheta, nr_target=args.nr)
        cb = cpu_reference_rate(args, src, anel)
        line = {"impl": "reference", "metric": "GLL-point updates/s", "value": cb["value"],
                "unit": "GLL-point updates/s", "n_gpus": N, "steps": args.steps,
                "warmup": args.warmup,
                # time one step of the named workload takes at the rate measured on the sample
                "ms_per_step": 25.0 * spec.nelem / cb["value"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "elements": int(spec.nelem),
                           "gll_points": int(25 * spec.nelem),
                           "sample_ms_per_step": cb["ms_per_step"],
                           "note": "reference Fortran/MPI solver cannot be built here (no "
                                   "Fortran compiler); this is the repo's CPU restatement of "
                                   "it on all host cores, bounded sample of the workload"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GLL-point updates/s",
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # stdout carries the one JSON line and nothing else (NCCL/torch banners go to stderr)
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from axisem_b200 import solver

    assert world == N or (world == 1 and N == 1), f"--gpus {N} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")

    K, W, R = args.steps, max(args.warmup, 3), max(args.repeats, 1)
    niter = W + (R + 2) * K + 8
    spec = prem_mesh_spec(ntheta=args.ntheta, nr_target=args.nr)
    t0 = time.perf_counter()
    att = AttenuationModel(coarse_grained=False) if (anel and args.full_memvars) else None
    prob = build_problem(spec, sp, anel=anel, att=att, niter=niter, rank=rank, nranks=world)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    loop = solver.time_loop(prob, device=local)
    t_upload = time.perf_counter() - t0
    # a stream of our own (not the legacy default stream): the library replays the Newmark step
    # from a CUDA graph captured on the stream it launches on
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    loop.set_stream(stream.cuda_stream)
    nel_s, nel_f, num_rec = prob.mesh.nel_solid, prob.mesh.nel_fluid, prob.num_rec
    stf_all = prob.stf.copy()
    prob.solid = prob.fluid = prob.att = prob.pw_solid = prob.pw_fluid = None   # host copies are no longer needed

    # seeded N(0,1)*1e-6 initial fields, so nothing is identically zero
    rng = np.random.default_rng(1234 + rank)
    for f in ("disp", "velo"):
        shp = loop._field_shape(f)
        loop.set(f, (rng.standard_normal(shp, dtype=np.float32) * np.float32(1e-6)))

    if world > 1:
        from axisem_b200.dist import connect_ranks
        connect_ranks(loop, rank, world, group=gloo)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up -------------------------------------------------------------------
    loop.run(W, sync=False)
    sync_all()
    launches0 = loop.gpu_launches

    # ---- timed regions: R times exactly K steps, device resident, each bracketed by a
    # barrier + synchronize; the line reports the median region (and min / max)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    region_ms = []
    for r in range(R):
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        loop.run(K, sync=False)
        ev1.record(stream)
        sync_all()
        region_ms.append(max_over_ranks(ev0.elapsed_time(ev1)))
    launches = (loop.gpu_launches - launches0) // R
    clk = clocks.stop() if rank == 0 else None
    ms = float(np.median(region_ms))
    # per-kernel breakdown of a step and the average launch time of the dominant kernel: a
    # separate pass of K steps with CUDA events around every launch (direct launches instead of
    # the graph replay; the event records cost a few microseconds per launch)
    loop.profile(1)
    sync_all()
    loop.run(K, sync=False)
    sync_all()
    prof_ms, prof_n = loop.get_profile()
    loop.profile(False)
    if world > 1:
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())

    total_pts = 25 * spec.nelem
    value = total_pts * K / (ms * 1e-3)

    # ---- e2e: same metric through the C ABI with host buffers, per step ---------------
    e2e = None
    if not args.no_e2e:
        stf_host = torch.from_numpy(stf_all).pin_memory().numpy()
        out = np.zeros((1, num_rec, 3), dtype=np.float32)
        import ctypes as C
        sync_all()
        t0 = time.perf_counter()
        for k in range(K):
            it = loop.iter
            loop.set_stf_values(it, stf_host[it:it + 1])               # host -> device
            loop.run(1, sync=False)
            if num_rec:                                                # D2H (synchronises)
                loop.lib.check(loop.lib.fn["fetch_seismograms"](
                    loop.h, C.c_int32(loop.nseismo - 1), C.c_int32(1),
                    out.ctypes.data_as(C.POINTER(C.c_float))))
            else:
                loop.synchronize()
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_pts * K / dt, "unit": "GLL-point updates/s",
               "h2d_bytes_per_step": 4, "d2h_bytes_per_step": int(12 * num_rec),
               "note": "per step: axb_set_stf_values (pinned host -> device), axb_run(1), "
                       "axb_fetch_seismograms (device -> host, synchronising); model arrays are "
                       "uploaded once at set-up (host->device "
                       f"{t_upload:.1f} s, not in the timed region)"}
    loop.close()
    del loop

    # ---- correctness of what these N ranks compute: a mid-size mesh through the same library
    # and halo wiring, seismograms against the committed 1-rank oracle golden
    check = None
    if not args.no_check:
        check = correctness_check(rank, world, local, gloo)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (S_A, class 0) -----------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    full = anel and args.full_memvars
    b_anel = (B_ANEL_FULL[pole] if full else B_ANEL[pole]) if anel else 0.0
    bytes_pt_sa = 4.0 * (2 * NC[pole] + NCOEF[pole]) + b_anel
    bytes_sa = bytes_pt_sa * 25 * nel_s
    # full memory variables: S_A is followed by k_anel_full (same profile class); one "launch"
    # below is then the pair
    ms_sa = prof_ms[0] / max(prof_n[0], 1) * (2 if full else 1)
    achieved = bytes_sa / (ms_sa * 1e-3) / 1e9 if ms_sa > 0 else None
    traffic = traffic_src = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        key = f"{pole}_{('anel_full' if full else 'anel') if anel else 'elastic'}"
        if key in tr:
            traffic = tr[key]["bytes_per_solid_element"] * nel_s
            traffic_src = tr[key]["source"]
    except Exception:
        pass
    step_bytes = 25.0 * ((B_SOLID[pole] + b_anel) * nel_s * world + B_FLUID[pole] * nel_f * world)
    names = ["solid_element(S_A)", "fluid_element(F_A)", "fluid_corrector(F_B)", "sf_coupling",
             "solid_corrector(S_B)", "halo", "sampling", "other"]
    line = {
        "metric": "GLL-point updates/s", "value": value, "unit": "GLL-point updates/s",
        "n_gpus": N, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload + f"; {100.0 * nel_f / (nel_s + nel_f):.1f} % of the elements are fluid "
                               "(outer core; 48 B/point against 227 for the anelastic solid)",
                   "elements": int(spec.nelem),
                   "gll_points": int(total_pts), "elements_per_gpu": int(spec.nelem // world),
                   "parallelism": f"theta-slices x{world}",
                   "l2": "inputs larger than L2: %.1f GB of algorithmic traffic per GPU per step"
                         % (step_bytes / world / 1e9),
                   "host_build_s": round(t_build, 1), "host_to_device_setup_s": round(t_upload, 1)},
        "repeats": {"n": R, "steps_each": K, "ms_per_step_median": ms / K,
                    "ms_per_step_min": min(region_ms) / K, "ms_per_step_max": max(region_ms) / K,
                    "spread_pct": 100.0 * (max(region_ms) - min(region_ms)) / ms},
        "clocks": clk,
        "e2e": e2e,
        "check": check,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": f"k_solid_tile<{pole}> (S_A)" + (" + k_anel_full" if full else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else
                                    "fallback 6650 GB/s (of fallback)",
                     "algorithmic_bytes_per_launch": bytes_sa, "avg_ms_per_launch": ms_sa,
                     "timing": "CUDA events around every S_A launch of a K-step pass next to the timed regions",
                     "step_algorithmic_GBs": step_bytes / world / (ms / K * 1e-3) / 1e9,
                     "step_frac_of_peak": step_bytes / world / (ms / K * 1e-3) / 1e9 / peak,
                     "kernel_ms_per_step": {n: prof_ms[i] / K for i, n in enumerate(names)
                                            if prof_n[i]}},
    }
    if N == 1 and not args.no_cpu_baseline:
        del prob
        cb = cpu_reference_rate(args, src, anel)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "same_mesh")}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


CHECK_GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_check_mtr_cg4.npz")


def check_problem(rank, world):
    """The mid-size case of the correctness check (tests/golden/make_bench_check.py made the
    golden from the same function with the oracle on one rank)."""
    z = np.load(CHECK_GOLDEN)
    spec = prem_mesh_spec(ntheta=int(z["ntheta"]), nr_target=int(z["nr"]), r_min_km=float(z["r_min_km"]))
    sp = SourceParams(src_type2="mtr", t_0=float(z["t_0"]))
    return build_problem(spec, sp, anel=True, niter=int(z["niter"]), rank=rank, nranks=world,
                         rec_colat_deg=z["colat_deg"], seis_it=int(z["seis_it"])), z


def correctness_check(rank, world, local, gloo):
    """Seismograms of the N-rank run (product library, same halo wiring as the timed run) against
    the 1-rank oracle golden: relative L2 over all stations and components."""
    import torch.distributed as dist
    from axisem_b200 import solver
    prob, z = check_problem(rank, world)
    loop = solver.time_loop(prob, device=local)
    if world > 1:
        from axisem_b200.dist import connect_ranks
        connect_ranks(loop, rank, world, group=gloo)
    loop.run(prob.niter)
    mine = (prob.rec_index, loop.seismograms())
    loop.close()
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=gloo)
    if rank != 0:
        return None
    ref = z["seismograms"].astype(np.float64)
    got = np.zeros_like(ref)
    for idx, s in parts:
        if len(idx):
            got[:, idx, :] = s
    err = float(np.sqrt(((got - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum()))
    return {"rel_l2": err, "tolerance": 1e-5, "ok": bool(err <= 1e-5), "ranks": world,
            "case": f"{int(z['ntheta'])}x{int(z['nr'])} PREM mesh, dipole, cg4 attenuation, "
                    f"{int(z['niter'])} Newmark steps, {ref.shape[1]} stations",
            "golden": "tests/golden/bench_check_mtr_cg4.npz (oracle, 1 rank)"}


if __name__ == "__main__":
    main()
