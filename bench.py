#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path.

metric : orientation x field evaluations per second through eigh + evolve (BASELINE.json).
step   : one pass of the path over the whole configuration table of the workload (default:
         BASELINE.json configs[4] / north_star target: mu + e + 3 1H + 14N, d = 96, 20 000
         orientations x 1 000 time points, T = inf) ending in the powder-averaged signal.
value  : configurations all ranks processed / max-over-ranks device time, inputs resident in HBM.
e2e    : same through the host-pointer C ABI call (musim_run_host): the configuration table is
         copied host->device and the result device->host inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4|c1] [--impl reference]

Under torchrun (N > 1) each rank owns one GPU and a round-robin shard of the orientation table
(experiment.py:369); the only collective is the final all-reduce of the [nt] signal.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# one BLAS thread per process: the CPU arms parallelise over worker processes (one per core),
# exactly like `mpirun -n <cores> muspinsim.mpi`; must be set before numpy is imported
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "orientation x field evaluations/sec (eigh+evolve)"
UNIT = "evaluations/s"


# ------------------------------------------------------------------------------------------
def make_spec(args, world=1):
    from muspinsim_b200 import workloads

    n = args.n_orient
    w = args.workload
    mult = world if args.scaling == "weak" else 1
    if w == "c5":
        return workloads.c5_large(n_orient=(n or 20000) * mult, nt=args.nt or 1000,
                                  temperature=np.inf if not args.general else 1.0)
    if w == "c2":
        return workloads.c2_hfine_powder(n_orient=(n or 20000) * mult, nt=args.nt or 1000,
                                         temperature=np.inf if not args.general else 1.0)
    if w == "c3":
        return workloads.c3_alc(n_orient=(n or 5000) * mult, n_field=args.nt or 2000)
    if w == "c4":
        return workloads.c4_fmuf_dissipation(n_orient=(n or 10000) * mult, nt=args.nt or 1000)
    if w == "c1":
        return workloads.c1_hfine()
    raise SystemExit("unknown workload " + w)


def algorithmic_flops(d, nt, mode):
    """real FP64 flops per evaluation."""
    F_eigh, F_gemm = 16.0 * d**3, 8.0 * d**3
    if mode == "fast":
        return F_eigh + F_gemm + 10.0 * nt * d * (d - 1) / 2
    if mode == "general":
        return F_eigh + 4 * F_gemm + 10.0 * nt * d * (d - 1) / 2
    if mode == "integral":
        return F_eigh + 4 * F_gemm + 12.0 * d * d
    n = d * d
    return 100.0 * n**3 + (8.0 / 3.0) * n**3 + 10.0 * nt * n


# per-kernel algorithmic flops per evaluation ("kernels and their rooflines")
def kernel_flops(name, d, nt):
    npairs = d * (d - 1) / 2
    return {
        "eigh_tridiag": (16.0 / 3) * d**3,  # zhetrd
        "eigh_tql": 30.0 * 1.2 * d * d,
        "eigh_apply": 6.0 * 1.2 * d**3,  # real Givens on d rows, ~1.2 d^2 rotations
        "eigh_back": (16.0 / 3) * d**3,  # zunmtr: reflectors applied to the real eigenvector matrix
        "eigh_jacobi": 16.0 * d**3,
        "rotate": 8.0 * d**3,  # one complex GEMM (fast path; O U is formed on the fly)
        # EXECUTED flops of the time-factorised kernel: 2 FMA per (pair incl. diagonal, time).  SURVEY 8(d)
        # counts 10 flops per (pair, time) for the direct cos/sin evaluation; that figure stays in the
        # whole-path number (algorithmic_flops) -- here it would put the kernel above the DFMA peak.
        "polar": 4.0 * nt * (npairs + d),
    }.get(name, 0.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (or the real reference when oracle/_ref travelled) on host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(payload):
    kind, spec, idx = payload
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    if kind == "reference":
        from oracle import ref_driver

        runner = ref_driver.make_runner(spec)
        cfg = runner.config
        t0 = time.perf_counter()
        for i in idx:
            snap = cfg[int(i)]
            cfg.store_time_slice(snap.id, runner.run_single(snap))
        return time.perf_counter() - t0, len(idx)
    from oracle import muspin_oracle as mo

    sys_ = mo.build_system(spec)
    cfg = mo.OracleConfig(spec)
    t0 = time.perf_counter()
    for i in idx:
        snap = cfg.snapshot(int(i))
        cfg.store_time_slice(snap["id"], mo.run_single(sys_, cfg, snap))
    return time.perf_counter() - t0, len(idx)


def cpu_baseline(spec, n_cfg, sample, procs=None):
    """Time `sample` configurations of `spec` split over `procs` worker processes, each taking
    cfg[r::P] like an MPI rank of the reference (mpi4py/mpirun are not installed)."""
    import multiprocessing as mp

    from oracle import muspin_oracle as mo
    from oracle import ref_driver

    mo.build_c() if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libfast_evolve.so")) else None
    kind = "reference" if ref_driver.available() else "port"
    procs = procs or os.cpu_count() or 1
    sample = min(sample, n_cfg)
    procs = min(procs, sample)
    idx = np.linspace(0, n_cfg - 1, sample).astype(int)
    parts = [idx[r::procs] for r in range(procs)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(kind, spec, p) for p in parts])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    return {
        "value": sample / busy,
        "unit": UNIT,
        "cores": procs,
        "kind": kind,
        "sample": "%d of %d configurations, %d worker processes each taking cfg[r::P] (mpi4py ranks: 0, not installed); "
                  "slowest worker %.1f s, wall %.1f s incl. start-up" % (sample, n_cfg, procs, busy, wall),
    }


def small_spec_for_cpu(args):
    """The reference builds a Python object per configuration while parsing; keep the .in
    small: the CPU arms run a bounded sample of the same system with fewer orientation rows."""
    a = argparse.Namespace(**vars(args))
    a.n_orient = args.cpu_sample
    a.scaling = "strong"
    return make_spec(a, 1)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = small_spec_for_cpu(args)
    from muspinsim_b200.configs import ConfigTable

    n_cfg = ConfigTable(spec).n_cfg
    vals = []
    for _ in range(args.warmup_ref + args.steps_ref):
        vals.append(cpu_baseline(spec, n_cfg, n_cfg))
    vals = vals[args.warmup_ref:]
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1], value=v)
    full = make_spec(args, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": 1e3 * n_cfg / v,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": full["name"], "sample_configurations": n_cfg},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--n-orient", type=int, default=0)
    ap.add_argument("--nt", type=int, default=0)
    ap.add_argument("--general", action="store_true", help="finite temperature (general evolve path)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="configurations in the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="library option key=value")
    args = ap.parse_args()
    if not args.cpu_sample:
        # bounded sample: ~5-10 s of work per host core
        per_core = {"c5": 32, "c2": 256, "c3": 8, "c4": 256, "c1": 1}[args.workload]
        args.cpu_sample = per_core * (os.cpu_count() or 1) if args.workload != "c1" else 1
    args.steps_ref, args.warmup_ref = max(1, min(args.steps, 2)), min(args.warmup, 1)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch

    from muspinsim_b200 import ExperimentRunner, _lib
    from muspinsim_b200.constants import MU_TAU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        from muspinsim_b200.dist import Communicator

        comm = Communicator(backend="nccl", device=local)

    spec = make_spec(args, world)
    runner = ExperimentRunner(spec, device=local)
    for kv in args.option:
        k, v = kv.split("=")
        runner.set_option(k, int(v))
    tab = runner.config
    d = runner.system.dim_total
    integral = tab.y == "integral"
    nt = 1 if integral else len(tab.times)
    sel = np.arange(tab.n_cfg)[rank::world]
    groups = runner._modes(sel)
    mode_name = {0: "general", 1: "fast", 2: "integral", 3: "lindblad", 4: "lindblad", 5: "integral"}[groups[0][0]]

    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    # device-resident inputs (for `value`) and pinned host copies (for `e2e`)
    dev_groups, host_groups = [], []
    for mode, idx in groups:
        order = idx[np.argsort(tab.slot[idx], kind="stable")]
        host = dict(B=np.ascontiguousarray(tab.B[order]), p=np.ascontiguousarray(tab.p[order]),
                    T=np.ascontiguousarray(tab.T[order]), w=np.ascontiguousarray(tab.w[order]),
                    slot=np.ascontiguousarray(tab.slot[order], dtype=np.int32))
        devt = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
        pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in host.items()}
        dev_groups.append((mode, len(order), devt))
        host_groups.append((mode, pinned))
    out_dev = torch.zeros(tab.n_slots, nt, dtype=torch.float64, device=dev)
    out_host = torch.zeros(tab.n_slots, nt, dtype=torch.float64).pin_memory().numpy()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    times = None if integral else tab.times
    handle = runner.handle

    def step_device():
        out_dev.zero_()
        for mode, n, t in dev_groups:
            handle.run_device(mode, n, t["B"].data_ptr(), t["p"].data_ptr(), t["T"].data_ptr(), t["w"].data_ptr(),
                              t["slot"].data_ptr(), times, MU_TAU, tab.n_slots, out_dev.data_ptr(), stream)
        if comm is not None:
            comm.sum_tensor_(out_dev)

    def step_host():
        out_host[...] = 0.0
        for mode, t in host_groups:
            handle.run_host(mode, t["B"], t["p"], t["T"], t["w"], t["slot"], times, MU_TAU, out_host)
        if comm is not None:
            out_host[...] = comm.sum_data(out_host)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()
        for a, b in ev:
            flush_buf.fill_(1)  # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if comm is not None:
            ms = comm.max_float(ms)
        return ms / steps

    warm = max(args.warmup, 3)
    # ---- device-resident timing, with per-kernel event timers on the same stream ----
    for _ in range(warm):
        step_device()
    torch.cuda.synchronize()
    handle.set_option("profile", 1)
    l0 = handle.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    launches = handle.launches - l0
    phases = {k: handle.phase_ms(k) / args.steps for k in
              ("eigh_tridiag", "eigh_tql", "eigh_apply", "eigh_back", "eigh_jacobi", "rotate", "rho0", "polar",
               "integral", "lindblad")}
    handle.set_option("profile", 0)
    result_dev = out_dev.cpu().numpy().copy()
    # ---- end to end through the host-pointer ABI ----
    ms_e2e = timed(step_host, args.steps, 2)
    if rank == 0:
        err = float(np.max(np.abs(result_dev - out_host)))
        assert err < 1e-9, "device-resident and host-pointer paths disagree: %g" % err

    n_total = tab.n_cfg
    n_local = len(sel)
    value = n_total / (ms_dev * 1e-3)
    e2e = n_total / (ms_e2e * 1e-3)
    h2d = sum(sum(v.nbytes for v in t.values()) for _, t in host_groups) + out_host.nbytes
    d2h = out_host.nbytes

    if rank != 0:
        if comm is not None:
            comm.close()
        return
    # ---- roofline of the dominant kernel ----
    peak_dfma = _lib.fp64_peak(local, 0)
    peak_dmma = _lib.fp64_peak(local, 1)
    top = max(phases, key=lambda k: phases[k])
    top_ms = phases[top]
    kf = kernel_flops(top, d, nt) * n_local
    achieved = kf / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
    path_flops = algorithmic_flops(d, nt, mode_name) * n_local
    # DRAM bytes per configuration of each phase at d = 96, from the ncu --set full capture in
    # profiles/r1_ncu_full_summary.md (dram__bytes_read.sum + dram__bytes_write.sum over 2960 configurations)
    ncu_dram_per_cfg_d96 = {"eigh_tridiag": 589.2e6 / 2960, "eigh_tql": 410.9e6 / 2960, "eigh_apply": 1570.4e6 / 2960,
                            "eigh_back": 821.9e6 / 2960, "rotate": 689.5e6 / 2960, "polar": 246.8e6 / 2960}
    traffic = ncu_dram_per_cfg_d96[top] * n_local if (d == 96 and top in ncu_dram_per_cfg_d96 and mode_name == "fast") else None
    roofline = {
        "bound": "fp64", "kernel": top, "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s",
        "frac": achieved / peak_dfma if peak_dfma else None, "traffic": traffic,
        "traffic_note": "bytes per step of the dominant phase (all its launches), ncu dram read+write per configuration x configurations",
        "peak_source": "measured live: DFMA micro-benchmark musim_fp64_peak (MEASURED_PEAKS.json has no FP64 entry); "
                       "DMMA m8n8k4 measured %.1f TFLOP/s" % peak_dmma,
        "kernel_ms_per_step": top_ms,
        "path": {"algorithmic_tflops": path_flops / (ms_dev * 1e-3) / 1e12,
                 "frac": path_flops / (ms_dev * 1e-3) / 1e12 / peak_dfma if peak_dfma else None},
    }
    cb = None
    if not args.no_cpu:
        try:
            sp = small_spec_for_cpu(args)
            from muspinsim_b200.configs import ConfigTable

            ncpu = ConfigTable(sp).n_cfg
            cb = cpu_baseline(sp, ncpu, ncpu)
        except Exception as exc:  # the baseline is a reported number, never a reason to fail
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "d": d, "configurations": n_total, "time_points": nt, "path": mode_name,
                   "l2": "256 MiB flush between timed iterations", "parallelism": "orientations sharded x%d" % world},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernel_ms_per_step": phases,
        "cpu_baseline": cb,
    }
    print(json.dumps(line))
    if comm is not None:
        comm.close()


if __name__ == "__main__":
    main()
