#!/usr/bin/env python
"""bench.py -- headline benchmark: batched acrobot iLQR solves/sec (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One STEP = one complete batched solve!: `batch` independent acrobot swing-up problems
(T=101, n=4, m=1, terminal goal constraint -> augmented-Lagrangian iLQR) warm-started from
the same synthetic initial guess each step and run to the reference's own termination.
Prints ONE JSON line (contract in the task statement):
  value      solves/sec with inputs resident in HBM (device-pointer entry points)
  e2e        solves/sec through the host-buffer C ABI: H2D of the initial guess, solve,
             D2H of trajectories + solver scalars, every step
  roofline   dominant kernel: algorithmic bytes / CUDA-event kernel time vs measured HBM peak
  cpu_baseline  the CPU oracle (oracle/ilqr_oracle.c, OpenMP, all host cores) on a bounded sample
Multi-GPU (torchrun, one rank per GPU): weak scaling, `batch` problems per GPU, no data-path
collective; one NCCL all_gather of trajectories + scalars per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_HORIZON = 101
METRIC = "ilqr_solves_per_sec_batched_acrobot_T101"
UNIT = "solves/s"


def synth_inputs(batch: int, T: int, seed: int = 0):
    """x1 = 0.1 N(0,1)^4 about the hanging equilibrium, u_t = N(0,1)."""
    rng = np.random.default_rng(seed)
    x1 = 0.1 * rng.standard_normal((batch, 4))
    ubar = rng.standard_normal((batch, T - 1, 1))
    return x1, ubar


def algorithmic_bytes(model, T: int, fused: bool = True) -> dict:
    """Per problem per tick, from SURVEY.md section 8d (doubles x 8 B; each array read / written once)."""
    n, m, p, cs, ct = model.n, model.m, model.p, model.cs, model.ct
    H = n * n + m * m + m * n
    lin_stage = 8 * ((n + m + p) + 3 * cs + H) + 8 * (n * n + n * m + n + m + H) + cs
    lin_term = 8 * (n + p + 3 * ct + n * n) + 8 * (n + n * n) + ct
    back_stage = 8 * (2 * n * n + 2 * n * m + m * m + n + m) + 8 * (m * n + 2 * m + n)
    back_term = 8 * (n * n + n)
    fwd_stage = 8 * (n + 2 * m + m * n + p) + 8 * (n + m + cs)
    fused_stage = 8 * ((n + m + p) + 3 * cs + 2 * H + m * n + 2 * m + n)   # fused K1+K2
    fused_term = 8 * (n + p + 3 * ct + 2 * n * n)
    if fused:
        return {"forward": (T - 1) * fwd_stage + 8 * (n + ct), "linearize": 0,
                "backward": (T - 1) * fused_stage + fused_term}
    return {
        "forward": (T - 1) * fwd_stage + 8 * (n + ct),
        "linearize": (T - 1) * lin_stage + lin_term,
        "backward": (T - 1) * back_stage + back_term,
    }


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(model, T, sample, steps, warmup, seed=0):
    """The reference's CPU path, restated (oracle/ilqr_oracle.c), all host threads, bounded sample."""
    from oracle.c_oracle import COracle

    x1, ubar = synth_inputs(sample, T, seed)
    co = COracle(model, T, sample, history_cap=1)
    xbar = co.rollout(x1, ubar)
    # every core this process may use -- not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or co.max_threads())
    times = []
    iters = 0
    for i in range(warmup + steps):
        co.initialize_controls(ubar)
        co.initialize_states(xbar)
        t0 = time.perf_counter()
        co.solve(cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            iters = int(co.get_stats()["iterations"].sum())
    total = sum(times)
    return {"value": sample * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample} problems of the same synthetic batch (seed {seed}), {len(times)} timed solve(s), "
                      f"C restatement of IterativeLQR.jl (Julia is not installable here), OpenMP dynamic schedule",
            "ms_per_step": 1e3 * total / len(times), "iterations_total": iters}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU and step")
    ap.add_argument("--slots", type=int, default=0,
                    help="solver slots the job is streamed through (default: 3 CTAs of 32 problems per B200 SM = 14208, at most steps x batch)")
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: W >= 3

    if args.slots <= 0:  # k_forward holds three 32-problem CTAs per SM (148 SMs), k_linback then takes three full waves
        args.slots = min(3 * 32 * 148, max(args.steps, 3) * args.batch)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import ilqr_b200  # noqa: F401
    from ilqr_b200 import problems

    model = problems.acrobot()
    T = T_HORIZON
    config = {"workload": f"acrobot swing-up, T={T}, n=4, m=1, terminal equality constraint (AL-iLQR), "
                          f"batch {args.batch} randomized initial states per GPU (BASELINE configs[1])",
              "step": "one batch of batch_per_gpu fresh problems per GPU; the K timed steps are submitted as one job and "
                      "streamed through `slots` solver slots (ilqr_solve_stream: a finished problem's slot is refilled "
                      "at once); lockstep_batches reports the same K batches as K separate ilqr_solve calls",
              "batch_per_gpu": args.batch, "slots": args.slots, "T": T, "options": "reference defaults (src/options.jl)",
              "l2": "per-tick working set (~1.5 GB at 14208 slots) exceeds the 126 MB L2; no flush"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(args.cpu_sample, args.batch)
        r = cpu_reference_run(model, T, sample, max(1, args.steps), min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ native arm
    import torch
    import torch.distributed as dist

    from ilqr_b200 import build, capi
    from ilqr_b200.distributed import gather_shards

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    K = args.steps
    n, m = model.n, model.m
    h = capi.Handle(build.model_library(model), T, n, m, model.p, model.cs, model.ct, B, device=local_rank, history_cap=8)
    hs = capi.Handle(build.model_library(model), T, n, m, model.p, model.cs, model.ct, args.slots, device=local_rank, history_cap=1)
    stream = torch.cuda.Stream(device=dev)
    h.set_stream(stream.cuda_stream)    # lock-step comparison: one batch per ilqr_solve
    hs.set_stream(stream.cuda_stream)   # the streamed job

    # synthetic job: K steps x B problems, every (rank, step) its own seed; nominal states by open-loop rollout
    nw = max(K, args.warmup)
    xs, us = [], []
    for step in range(nw):
        x1, ubar = synth_inputs(B, T, seed=1000 * rank + step)
        xs.append(h.rollout(x1, ubar)); us.append(ubar)
    hx = torch.from_numpy(np.concatenate(xs)).pin_memory()      # [nw*B][T][n]  pinned host (e2e)
    hu = torch.from_numpy(np.concatenate(us)).pin_memory()
    dx, du = hx.to(dev), hu.to(dev)                              # device-resident copies (value)
    NB = K * B
    NO = nw * B  # output buffers also hold the (possibly longer) warm-up job
    ox = torch.empty((NO, T, n), dtype=torch.float64, device=dev)
    ou = torch.empty((NO, T - 1, m), dtype=torch.float64, device=dev)
    oit = torch.zeros(NO, dtype=torch.int32, device=dev)
    ost = torch.zeros(NO, dtype=torch.uint8, device=dev)
    oJ = torch.zeros(NO, dtype=torch.float64, device=dev)
    omv = torch.zeros(NO, dtype=torch.float64, device=dev)
    out_hx = torch.empty((NB, T, n), dtype=torch.float64).pin_memory()
    out_hu = torch.empty((NB, T - 1, m), dtype=torch.float64).pin_memory()

    def gather():
        if world > 1:
            sc = torch.stack([oit[:NB].double(), ost[:NB].double(), oJ[:NB], omv[:NB]], dim=1)
            with torch.cuda.stream(stream):
                gather_shards({"x": ox[:NB], "u": ou[:NB], "scalars": sc}, NB * world, dist)

    def job_resident(nsteps=K):
        """the engine's production mode: nsteps*B fresh problems streamed through the solver's slots (continuous batching)"""
        hs.solve_stream(nsteps * B, dx.data_ptr(), du.data_ptr(), 0, ox.data_ptr(), ou.data_ptr(), oit.data_ptr(),
                       ost.data_ptr(), oJ.data_ptr(), omv.data_ptr(), 0, 0)
        if nsteps == K:
            gather()

    def job_e2e():
        return hs.solve_stream_host(hx.numpy()[:NB], hu.numpy()[:NB], out_x=out_hx.numpy(), out_u=out_hu.numpy())

    def steps_lockstep():
        """K separate ilqr_solve calls, one batch each (every batch waits for its slowest problem)"""
        for step in range(K):
            h.initialize_controls_device(du[step * B:(step + 1) * B].data_ptr())
            h.initialize_states_device(dx[step * B:(step + 1) * B].data_ptr())
            h.solve()
            h.get_trajectory_device(ox.data_ptr(), ou.data_ptr())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # warm-up: W steps' worth of problems through both paths (graph capture, allocator, clocks)
    job_resident(args.warmup)
    steps_lockstep()
    job_e2e()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    launches0 = int(hs.get_counters()["launches"])
    ms_value = timed(job_resident)                       # headline: K*B problems, inputs resident in HBM
    launches_timed = int(hs.get_counters()["launches"]) - launches0
    iters = oit[:NB].cpu().numpy().copy()
    viol = omv[:NB].cpu().numpy().copy()
    ms_e2e = timed(job_e2e)                              # same job through the host-buffer C ABI call
    ms_lock = timed(steps_lockstep)                      # K lock-step batch solves (ilqr_solve), for comparison
    hs.set_profiling(True)                               # the K*B-problem job again with CUDA events around every kernel
    ms_prof = timed(job_resident)                        # (set_profiling(True) zeroes the counters: they cover this pass)
    counters = hs.get_counters()
    hs.set_profiling(False)
    clock_info = clocks.stop() if clocks else None

    total = NB * world
    value = total / (ms_value * 1e-3)
    e2e_value = total / (ms_e2e * 1e-3)
    h2d = (NB * T * n + NB * (T - 1) * m) * 8
    d2h = h2d + NB * (4 + 1 + 8 + 8 + 8 + 4)
    stats = {"iterations": iters, "max_violation": viol}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (largest share of the timed step)
    names = ["forward", "linearize", "backward"]
    ab = algorithmic_bytes(model, T)
    kms = [float(v) for v in counters["kernel_ms"]]
    kl = [int(v) for v in counters["kernel_launches"]]
    pt = int(counters["problem_ticks"])
    peak, peak_src = hbm_peak()
    kernels = {}
    for i, nm in enumerate(names):
        gbs = ab[nm] * pt / (kms[i] * 1e-3) / 1e9 if kms[i] > 0 else 0.0
        kernels[nm] = {"ms_total": kms[i], "launches": kl[i], "us_per_launch": 1e3 * kms[i] / max(kl[i], 1),
                       "share_of_step": kms[i] / ms_prof, "algorithmic_bytes_per_problem_tick": ab[nm],
                       "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    dom = max(names, key=lambda nm: kernels[nm]["ms_total"])
    ticks = int(counters["ticks"])
    kname = {"forward": "k_forward", "linearize": "k_linearize", "backward": "k_linback (fused gradients!+backward_pass!)"}
    traffic, traffic_n = None, 4096
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/summarize.py)
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_launch"].get({"forward": "k_forward", "backward": "k_linback", "linearize": "k_linearize"}[dom])
        traffic_n = int(tj.get("problems_per_launch", 4096))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic,
                "traffic_note": f"bytes per launch with all {traffic_n} problems iterating (ncu --set full, profiles/r1_traffic.json); "
                                f"algorithmic bytes per such launch = algorithmic_bytes_per_problem_tick x {traffic_n}",
                "peak_source": peak_src,
                "how": "algorithmic bytes per problem-tick (SURVEY 8d) x problem-ticks / sum of that kernel's CUDA-event "
                       "durations on the solve stream, over a second pass of the same K-step job with events around every "
                       "kernel (ms_per_step_with_kernel_events); the headline pass runs the same kernels from CUDA graphs. "
                       "The fused k_linback (gradients! + backward_pass!) is listed under 'backward'; its algorithmic bytes "
                       "are the fused figure of SURVEY 8d",
                "ms_per_step_with_kernel_events": ms_prof / args.steps,
                "kernels": kernels}

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(model, T, min(args.cpu_sample, B), 1, 0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "ms_per_iteration": ms_prof / max(ticks, 1),
            "ticks_per_step": ticks / args.steps,
            "lockstep_batches": {"value": total / (ms_lock * 1e-3), "unit": UNIT, "ms_per_step": ms_lock / args.steps,
                                 "what": "K separate ilqr_solve calls of one batch each: every batch waits for its slowest problem"},
            "iterations_per_problem": {"mean": float(stats["iterations"].mean()), "max": int(stats["iterations"].max())},
            "converged_frac": float((stats["max_violation"] <= 5e-3).mean()),
            "problems_per_step_per_gpu": B,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d // K), "d2h_bytes_per_step": int(d2h // K),
                    "ms_per_step": ms_e2e / args.steps, "call": "ilqr_solve_stream_host (pinned host buffers in, host buffers out)"},
            "gpu_launches": launches_timed,
            "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
