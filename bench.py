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
(experiment.py:369); the only collective is the final all-reduce of the [nt] signal.  The default
is STRONG scaling -- the north_star's fixed table (20 000 orientations) split over the N GPUs -- and
the line also carries the weak-scaling figure (20 000 orientations per GPU) as `weak`.

parity : the line's `parity` block compares the GPU path, configuration by configuration, with the
         results the CPU arm (the unmodified reference when oracle/_ref travelled, else the oracle
         port) produced for its timed sample -- same system, same indices, one output row per
         configuration; the run FAILS above 1e-9.
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
        "eigh_tdc": (8.0 / 3) * d**3,  # dstedc: leaves + two merge levels, eigenvector GEMMs (real)
        "eigh_apply": 6.0 * 1.2 * d**3,  # real Givens on d rows, ~1.2 d^2 rotations
        "eigh_back": (16.0 / 3) * d**3,  # zunmtr: reflectors applied to the real eigenvector matrix
        "eigh_jacobi": 16.0 * d**3,
        "rotate": 8.0 * d**3,  # one complex GEMM (fast path; O U is formed on the fly)
        # EXECUTED flops of the time-factorised kernel: 2 FMA per (pair incl. diagonal, time).  SURVEY 8(d)
        # counts 10 flops per (pair, time) for the direct cos/sin evaluation; that figure stays in the
        # whole-path number (algorithmic_flops) -- here it would put the kernel above the DFMA peak.
        "polar": 4.0 * nt * (npairs + d),
    }.get(name, 0.0)


# phase -> kernel-name pattern, for the DRAM traffic read from the tracked ncu summary (profiles/*.json,
# written by tools/ncu_full_summary.py from an `ncu --set full` capture of this same command)
PHASE_KERNELS = {"eigh_tridiag": "hql_tridiag", "eigh_tql": "hql_tql", "eigh_tdc": "tdc_", "eigh_apply": "hql_apply", "eigh_back": "hql_backwy|hql_reflect|hql_tfactor",
                 "rotate": "zgemm_dmma", "polar": "polar_", "lindblad": "lind_|zgemm_dmma|cgemm_", "eigh_jacobi": "eigh_jacobi"}


def ncu_traffic(workload, general, phase):
    """DRAM bytes per configuration of `phase` from the tracked ncu summary, or (None, reason)."""
    import re

    name = "r2_ncu_%s%s.json" % (workload, "_general" if general else "")
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None, "no profiles/%s" % name
    doc = json.load(open(path))
    pat = re.compile(PHASE_KERNELS.get(phase, "$^"))
    tot = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in doc["kernels"] if pat.search(k["name"]))
    if tot <= 0 or not doc.get("configurations"):
        return None, "phase not in profiles/%s" % name
    return tot / doc["configurations"], "profiles/%s (%d configurations per launch group in the capture)" % (name, doc["configurations"])


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
        vals = []
        t0 = time.perf_counter()
        for i in idx:
            snap = cfg[int(i)]
            data = runner.run_single(snap)  # experiment.py:434-498: the weighted real signal of ONE configuration
            cfg.store_time_slice(snap.id, data)
            vals.append(np.atleast_1d(np.array(data, dtype=float)))
        return time.perf_counter() - t0, len(idx), idx, vals
    from oracle import muspin_oracle as mo

    sys_ = mo.build_system(spec)
    cfg = mo.OracleConfig(spec)
    vals = []
    t0 = time.perf_counter()
    for i in idx:
        snap = cfg.snapshot(int(i))
        data = mo.run_single(sys_, cfg, snap)
        cfg.store_time_slice(snap["id"], data)
        vals.append(np.atleast_1d(np.array(data, dtype=float)))
    return time.perf_counter() - t0, len(idx), idx, vals


def cpu_baseline(spec, n_cfg, sample, procs=None, keep=None):
    """Time `sample` configurations of `spec` split over `procs` worker processes, each taking
    cfg[r::P] like an MPI rank of the reference (mpi4py/mpirun are not installed).  With
    keep = {} the per-configuration results come back as keep["idx"], keep["vals"] for the
    parity check."""
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
    if keep is not None:
        keep["idx"] = np.concatenate([r[2] for r in res])
        keep["vals"] = np.array([v for r in res for v in r[3]])
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
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
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
    from muspinsim_b200.configs import ConfigTable
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
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    warm = max(args.warmup, 3)
    PHASES = ("eigh_tridiag", "eigh_tdc", "eigh_tql", "eigh_apply", "eigh_back", "eigh_jacobi", "rotate", "rho0", "polar",
              "integral", "lindblad")

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

    def measure(spec, steps, with_e2e, sample_clocks):
        """One workload on this rank's shard cfg[rank::world]: device-resident timing with per-kernel
        event timers, optionally the host-pointer (e2e) timing.  Returns a dict."""
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

        # ---- device-resident timing, with per-kernel event timers on the same stream ----
        for _ in range(warm):
            step_device()
        torch.cuda.synchronize()
        handle.set_option("profile", 1)
        l0 = handle.launches
        g0 = handle.phase_ms("lind_gemm_cfgs")
        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()
        ms_dev = timed(step_device, steps, 0)
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        res = dict(runner=runner, tab=tab, d=d, nt=nt, mode_name=mode_name, n_total=tab.n_cfg, n_local=len(sel),
                   ms_dev=ms_dev, clocks=clocks, launches=handle.launches - l0,
                   lind_gemm_cfgs=(handle.phase_ms("lind_gemm_cfgs") - g0) / steps,
                   phases={k: handle.phase_ms(k) / steps for k in PHASES}, spec=spec)
        handle.set_option("profile", 0)
        if with_e2e:
            result_dev = out_dev.cpu().numpy().copy()
            # ---- end to end through the host-pointer ABI ----
            res["ms_e2e"] = timed(step_host, steps, 2)
            if rank == 0:
                err = float(np.max(np.abs(result_dev - out_host)))
                assert err < 1e-9, "device-resident and host-pointer paths disagree: %g" % err
            res["h2d"] = sum(sum(v.nbytes for v in t.values()) for _, t in host_groups) + out_host.nbytes
            res["d2h"] = out_host.nbytes
        return res

    spec = make_spec(args, world)
    R = measure(spec, args.steps, True, True)
    runner, tab, d, nt, mode_name = R["runner"], R["tab"], R["d"], R["nt"], R["mode_name"]
    handle = runner.handle
    ms_dev, ms_e2e, phases = R["ms_dev"], R["ms_e2e"], R["phases"]
    n_total, n_local = R["n_total"], R["n_local"]
    value = n_total / (ms_dev * 1e-3)
    e2e = n_total / (ms_e2e * 1e-3)

    # ---- the other scaling mode as a second field (N > 1 only) ----
    other = None
    if world > 1 and args.workload != "c1":
        a2 = argparse.Namespace(**vars(args))
        a2.scaling = "weak" if args.scaling == "strong" else "strong"
        del R["runner"]
        runner = None
        O = measure(make_spec(a2, world), max(2, args.steps // 2), False, False)
        other = {"scaling": a2.scaling, "value": O["n_total"] / (O["ms_dev"] * 1e-3), "unit": UNIT,
                 "ms_per_step": O["ms_dev"], "configurations": O["n_total"], "kernel_ms_per_step": O["phases"]}
        handle = O["runner"].handle  # same system: serves the parity check below

    if rank != 0:
        if comm is not None:
            comm.close()
        return

    # ---- whole-API wall time: ExperimentRunner(spec).run() incl. table build, handle creation,
    #      workspace allocation, device-side table expansion and the result copy (one GPU) ----
    api = None
    if world == 1:
        walls = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r2 = ExperimentRunner(spec, device=local)
            r2.run()
            walls.append(time.perf_counter() - t0)
            r2.handle.close()
        api = {"value": n_total / min(walls), "unit": UNIT, "wall_s": min(walls), "first_call_wall_s": walls[0],
               "what": "ExperimentRunner(spec).run(): spec -> system matrices, configuration table, new device handle, "
                       "workspaces (the second call re-uses the blocks the first call's destroyed handle left in the library's "
                       "cache; first_call_wall_s pays cudaMalloc), evaluation, result on the host (best of 2)"}

    # ---- roofline of the dominant kernel ----
    peak_dfma = _lib.fp64_peak(local, 0)
    peak_dmma = _lib.fp64_peak(local, 1)
    top = max(phases, key=lambda k: phases[k])
    top_ms = phases[top]
    if top == "lindblad":
        # EXECUTED flops of the expm formulation (the reference's zgeev count does not apply): the
        # batched n x n complex GEMMs actually launched (scaling and squaring: counted by the library)
        # + the series kernel's (NA + NB) mat-vecs and nt dot products per configuration
        n = d * d
        NB = 1
        while NB * NB < nt and NB < 16:
            NB <<= 1
        kf = 8.0 * n**3 * R["lind_gemm_cfgs"] + n_local * (8.0 * n * n * (NB + (nt + NB - 1) // NB) + 8.0 * n * nt)
    else:
        kf = kernel_flops(top, d, nt) * n_local
    achieved = kf / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
    path_flops = algorithmic_flops(d, nt, mode_name) * n_local
    per_cfg, tsrc = ncu_traffic(args.workload, args.general, top)
    roofline = {
        "bound": "fp64", "kernel": top, "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s",
        "frac": achieved / peak_dfma if peak_dfma else None,
        "traffic": per_cfg * n_local if per_cfg else None,
        "traffic_note": "DRAM read+write bytes per step of the dominant phase (all its launches): " + tsrc,
        "flops_model": "executed (expm GEMMs counted by the library + series kernel)" if top == "lindblad" else "algorithmic (DESIGN.md section 3)",
        "peak_source": "measured live: DFMA micro-benchmark musim_fp64_peak (MEASURED_PEAKS.json has no FP64 entry); "
                       "DMMA m8n8k4 measured %.1f TFLOP/s" % peak_dmma,
        "kernel_ms_per_step": top_ms,
        "path": {"algorithmic_tflops": path_flops / (ms_dev * 1e-3) / 1e12,
                 "frac": path_flops / (ms_dev * 1e-3) / 1e12 / peak_dfma if peak_dfma else None,
                 "note": "SURVEY 8(d) flop count of the whole path (charges 10 flops per (pair, time) which the NUFFT kernel does not execute)",
                 # the honest whole-path figure: the same count with the time loop at what the type-1 NUFFT executes
                 # (~150 flops per pair, independent of nt) when that kernel ran
                 "executed_frac": ((path_flops - (10.0 * nt - 150.0) * d * (d - 1) / 2 * n_local) / (ms_dev * 1e-3) / 1e12 / peak_dfma
                                   if (peak_dfma and mode_name in ("fast", "general") and nt >= 96) else None)},
    }
    # ---- CPU arm on a bounded sample + per-configuration parity of the GPU path on the same sample ----
    cb, parity = None, None
    if not args.no_cpu:
        try:
            sp = small_spec_for_cpu(args)
            stab = ConfigTable(sp)
            keep = {}
            cb = cpu_baseline(sp, stab.n_cfg, stab.n_cfg, keep=keep)
        except Exception as exc:  # the baseline is a reported number, never a reason to fail
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
            keep = None
        if keep:
            # GPU: the same configurations, one output row each, weights as run_single applies them
            idx = keep["idx"]
            small = ExperimentRunner(sp, device=local)
            small._handle = handle  # same spin system: reuse the resident handle
            got = np.zeros((len(idx), nt))
            for mode, sub in small._modes(np.arange(stab.n_cfg)):
                pick = np.nonzero(np.isin(idx, sub))[0]
                ii = idx[pick]
                part = np.zeros((len(ii), nt))
                handle.run_host(mode, stab.B[ii], stab.p[ii], stab.T[ii], stab.w[ii] * stab.avg_N,
                                np.arange(len(ii)), None if stab.y == "integral" else stab.times, MU_TAU, part)
                got[pick] = part
            err = float(np.max(np.abs(got - keep["vals"].reshape(len(idx), -1))))
            parity = {"max_abs_err": err, "n": int(len(idx)), "tol": 1e-9, "against": cb["kind"],
                      "what": "per-configuration signal of the CPU arm's timed sample vs musim_run_host on the same "
                              "configurations (one output row each), full time grid"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "d": d, "configurations": n_total, "time_points": nt, "path": mode_name,
                   "l2": "256 MiB flush between timed iterations", "parallelism": "orientations sharded x%d" % world},
        "clocks": R["clocks"],
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(R["h2d"]),
                "d2h_bytes_per_step": int(R["d2h"])},
        "api_e2e": api,
        "gpu_launches": int(R["launches"]),
        "roofline": roofline,
        "kernel_ms_per_step": phases,
        "parity": parity,
        "cpu_baseline": cb,
    }
    if other is not None:
        line[other["scaling"]] = other
    print(json.dumps(line))
    if comm is not None:
        comm.close()
    if parity is not None and not (parity["max_abs_err"] < parity["tol"]):
        raise SystemExit("parity FAILED: max |gpu - %s| = %.3e over %d configurations" % (parity["against"], parity["max_abs_err"], parity["n"]))


if __name__ == "__main__":
    main()
