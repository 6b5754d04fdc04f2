#!/usr/bin/env python
"""bench.py -- batched iLQR solves/sec on B200 (BASELINE.json's metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config c2|c3|c4|c5]

Configs (BASELINE.json `configs`; c2 is the one the metric is quoted on and the default):
  c2  acrobot swing-up, T=101, n=4, m=1, terminal equality (AL-iLQR), 4096 randomized problems per GPU and step
  c3  car with obstacle / box inequality constraints, T=51, n=3, m=2, 2048 problems per GPU and step (16384 over 8 GPUs)
  c4  synthetic dense linear-quadratic tracking, T=256, n=64, m=16, batch 1024 (wide-model Riccati path)
  c5  receding-horizon MPC: acrobot T=101, 1024 problems per GPU (8192 over 8), --steps warm-started re-solves each

One STEP = one pass of the hot path over one batch: `batch` complete solve! calls (c2-c4) or one receding-horizon
re-solve of every problem (c5).  c2/c3 submit the K timed steps as ONE streamed job (ilqr_solve_stream: continuous
batching through `slots` solver slots); `lockstep_batches` reports the same K batches as K separate ilqr_solve calls.
  value         units/s with inputs resident in HBM (device-pointer entry points), whole job over all ranks
  e2e           the same job through the host-buffer C ABI: H2D of the inputs, solve, D2H of the results
  roofline      dominant kernel: algorithmic bytes (or flops) / CUDA-event kernel time vs the measured peak
  cpu_baseline  the CPU oracle (oracle/ilqr_oracle.c, OpenMP, all host cores) on a bounded sample of the SAME inputs
  parity        the GPU results of that sample against the oracle's (iteration counts, trajectories)
Multi-GPU (torchrun, one rank per GPU): weak scaling, no data-path collective; one all-gather of trajectories and
solver scalars per job through the C ABI (ilqr_gather = ncclAllGather), timed inside `value` and reported on its own.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CONFIGS = {
    "c2": dict(model="acrobot", T=101, batch=4096, mode="stream", cpu_sample=1024, unit="solves/s",
               metric="ilqr_solves_per_sec_batched_acrobot_T101",
               workload="acrobot swing-up, T=101, n=4, m=1, terminal equality constraint (AL-iLQR), batch {batch} "
                        "randomized initial states per GPU (BASELINE configs[1])"),
    "c3": dict(model="car", T=51, batch=2048, mode="stream", cpu_sample=1024, unit="solves/s",
               metric="ilqr_solves_per_sec_batched_car_T51",
               workload="car with obstacle + box inequality constraints (AL-iLQR), T=51, n=3, m=2, c_s=5, c_T=4, batch "
                        "{batch} per GPU = 16384 sharded over 8 GPUs (BASELINE configs[2])"),
    "c4": dict(model="lq64", T=256, batch=1024, mode="lockstep", cpu_sample=16, unit="solves/s",
               metric="ilqr_solves_per_sec_dense_lq_n64_m16_T256",
               workload="synthetic dense linear-quadratic tracking, T=256, n=64, m=16, p=128, batch {batch} "
                        "(wide-model Riccati path, BASELINE configs[3])"),
    "c5": dict(model="acrobot", T=101, batch=1024, mode="mpc", cpu_sample=64, unit="re-solves/s",
               metric="ilqr_mpc_resolves_per_sec_acrobot_T101",
               workload="receding-horizon MPC: acrobot T=101, batch {batch} per GPU = 8192 over 8 GPUs, one warm-started "
                        "re-solve of every problem per step, closed loop on the model plant (BASELINE configs[4])"),
}


def synth_inputs(batch: int, T: int, seed: int = 0):
    """x1 = 0.1 N(0,1)^4 about the hanging equilibrium, u_t = N(0,1)."""
    rng = np.random.default_rng(seed)
    x1 = 0.1 * rng.standard_normal((batch, 4))
    ubar = rng.standard_normal((batch, T - 1, 1))
    return x1, ubar


def config_inputs(cfg, batch: int, seed: int):
    """(model, x1, ubar, w) of one batch: SURVEY.md 8d randomisation, numpy PCG64 so that oracle and engine see the same arrays."""
    from common import inputs, lq_inputs
    from ilqr_b200 import problems

    T = cfg["T"]
    if cfg["model"] == "acrobot":
        x1, ubar = synth_inputs(batch, T, seed)
        return problems.acrobot(), x1, ubar, None
    if cfg["model"] == "car":
        model, x1, ubar = inputs("car", batch, T, seed)
        return model, x1, ubar, None
    if cfg["model"] == "lq64":
        # the plant scaling s_b and the reference r_{b,t} (the per-step parameters w) belong to the problem slot: one draw per
        # rank (seed of its step 0), uploaded once; the initial state x1 is redrawn for every step
        model, _, ubar, w = lq_inputs(batch, T, 64, 16, seed=seed - seed % 1000)
        x1 = np.random.default_rng(seed + 7919).standard_normal((batch, 64))
        if os.environ.get("C4_MODEL", "") == "banded":
            model = problems.lq_banded(64, 16)
        ubar[:] = 0.0  # u = 0, x = rollout
        return model, x1, ubar, w
    raise KeyError(cfg["model"])


def job_seed(cfg, rank: int, step: int) -> int:
    """seed of the batch (rank, step) works on; SURVEY 8d: seed 0 for C2-C4, seed 2 for the MPC config"""
    return 1000 * rank + step + (2 if cfg["mode"] == "mpc" else 0)


def sample_inputs(cfg, sample: int, step: int):
    """the first `sample` problems of rank 0's batch of `step` (the draw depends on the batch size, so draw the whole batch)"""
    model, x1, ubar, w = config_inputs(cfg, cfg["batch"], job_seed(cfg, 0, step))
    return model, x1[:sample], ubar[:sample], (w[:sample] if w is not None else None)


def algorithmic_bytes(model, T: int, fused: bool = True) -> dict:
    """Per problem per tick, from SURVEY.md section 8d (doubles x 8 B; each array read / written once)."""
    n, m, p, cs, ct = model.n, model.m, model.p, model.cs, model.ct
    H = n * n + m * m + m * n
    lin_stage = 8 * ((n + m + p) + 3 * cs + H) + 8 * (n * n + n * m + n + m + H) + cs
    lin_term = 8 * (n + p + 3 * ct + n * n) + 8 * (n + n * n) + ct
    back_stage = 8 * (2 * n * n + 2 * n * m + m * m + n + m) + 8 * (m * n + 2 * m + n)
    back_term = 8 * (n * n + n)
    fwd_stage = 8 * (n + 2 * m + m * n + p) + 8 * (n + m + cs)
    # SURVEY 8d's K4 row leaves out what the expected-decrease sweep of forward_pass! reads (src/data/methods.jl:42-54,
    # src/forward_pass.jl:19-20: fx, fu and the Lagrangian gradient); reported next to the contract's figure
    fwd_dgp_stage = 8 * (n * n + n * m + n + m)
    fused_stage = 8 * ((n + m + p) + 3 * cs + 2 * H + m * n + 2 * m + n)   # fused K1+K2
    fused_term = 8 * (n + p + 3 * ct + 2 * n * n)
    if fused:
        return {"forward": (T - 1) * fwd_stage + 8 * (n + ct), "linearize": 0,
                "backward": (T - 1) * fused_stage + fused_term, "forward_dgp_inputs": (T - 1) * fwd_dgp_stage}
    return {"forward": (T - 1) * fwd_stage + 8 * (n + ct), "linearize": (T - 1) * lin_stage + lin_term,
            "backward": (T - 1) * back_stage + back_term, "forward_dgp_inputs": (T - 1) * fwd_dgp_stage}


def riccati_flops(model, T: int) -> float:
    """4n^3 + 10n^2 m + 6nm^2 + m^3/3 per (problem, time step)."""
    n, m = model.n, model.m
    return (4 * n**3 + 10 * n * n * m + 6 * n * m * m + m**3 / 3) * (T - 1)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """FP64 tensor-pipe (DMMA m8n8k4) throughput measured on this pool's B200s by profiles/microbench/fp64_peak.cu -- the pipe the
    wide-model Riccati kernel's contractions run on, and the higher of the two FP64 peaks (the driver's file has no FP64 entry)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_fp64_peak.json")) as f:
            return float(json.load(f)["dmma_m8n8k4_tflops"]), "measured (profiles/r2_fp64_peak.json: mma.sync.m8n8k4.f64 loop; DFMA loop 33.7)"
    except Exception:
        return 37.0, "fallback (spec)"


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


# ---------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_reference_run(cfg, sample: int, steps: int, warmup: int, keep_first: bool = False):
    """The reference's CPU path, restated (oracle/ilqr_oracle.c), all host threads, on the first `sample` problems of
    the SAME batches the GPU job of rank 0 solves (seed = step); `warmup` untimed + `steps` timed steps."""
    from oracle.c_oracle import COracle

    T, mode = cfg["T"], cfg["mode"]
    model, x1, ubar, w = sample_inputs(cfg, sample, 0)
    co = COracle(model, T, sample, history_cap=1)
    # every core this process may use -- not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or co.max_threads())
    times, iters, first = [], 0, None
    units = 0
    if mode == "mpc":
        xbar = co.rollout(x1, ubar)
        co.initialize_controls(ubar); co.initialize_states(xbar); co.solve(cores)   # the initial solve (untimed, like the GPU arm)
        first = dict(applied_u=[], x_next=[])
        for i in range(warmup + steps):
            before = co.get_stats()["iterations"].astype(np.int64).sum()
            t0 = time.perf_counter()
            au, xn = co.mpc_step(cores)
            dt = time.perf_counter() - t0
            if keep_first and len(first["applied_u"]) < 8:
                first["applied_u"].append(au.copy()); first["x_next"].append(xn.copy())
            if i >= warmup:
                times.append(dt); units += sample
                iters += int(co.get_stats()["iterations"].astype(np.int64).sum() - (0 if model.constrained else before))
    else:
        for i in range(warmup + steps):
            step = max(i - warmup, 0)  # warm-up re-uses step 0's batch
            _, x1, ubar, w = sample_inputs(cfg, sample, step)
            co = COracle(model, T, sample, history_cap=1)  # a FRESH solver per step, like every problem of the GPU job (a re-used
                                                           # solver carries its current trajectory and constraint values over, Q2)
            if w is not None:
                co.set_parameters(w)
            xbar = co.rollout(x1, ubar)
            co.initialize_controls(ubar); co.initialize_states(xbar)
            t0 = time.perf_counter()
            co.solve(cores)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt); units += sample
                iters += int(co.get_stats()["iterations"].sum())
            if keep_first and i == 0:
                xo, uo = co.get_trajectory()
                st = co.get_stats()
                first = dict(x=xo, u=uo, iterations=st["iterations"].copy(), objective=st["objective"].copy())
    total = sum(times)
    return {"value": units / total, "unit": cfg["unit"], "cores": cores, "kind": "port",
            "sample": f"first {sample} problems of rank 0's batches (steps 0..{steps - 1}, the GPU job's own inputs), {warmup} warm-up + "
                      f"{len(times)} timed step(s), C restatement of IterativeLQR.jl (Julia is not installable here), OpenMP dynamic schedule",
            "ms_per_step": 1e3 * total / len(times), "iterations_per_unit": iters / max(units, 1), "first": first}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU and step (default: the config's)")
    ap.add_argument("--slots", type=int, default=0, help="solver slots a streamed job runs through (default: see DEFAULT_SLOTS)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: W >= 3
    cfg = dict(CONFIGS[args.config])
    if args.batch > 0:
        cfg["batch"] = args.batch
    B, T, K, mode = cfg["batch"], cfg["T"], args.steps, cfg["mode"]
    sample = min(args.cpu_sample or cfg["cpu_sample"], B)
    METRIC, UNIT = cfg["metric"], cfg["unit"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import ilqr_b200  # noqa: F401

    if mode == "stream" and args.slots <= 0:
        # measured on the 20-step jobs (profiles/README.md): 8 x 148 x 32 = 37888 slots -- one wave of the thread-per-problem
        # backward kernel at 8 warps per SM -- for both models (acrobot 126-129 k solves/s against 103 k at 14208 and 107 / 125 k
        # at 47360 / 56832; car 542 k against 375 k at 9472 and 485 / 503 k at 47360 / 56832)
        args.slots = min(8 * 32 * 148, max(K, 3) * B)
    config = {"workload": cfg["workload"].format(batch=B), "config": args.config,
              "step": {"stream": "one batch of batch_per_gpu fresh problems per GPU; the K timed steps are submitted as one job and streamed "
                                 "through `slots` solver slots (ilqr_solve_stream: a finished problem's slot is refilled at once, the drain "
                                 "phase packs the running problems into fewer blocks); lockstep_batches = the same K batches as K ilqr_solve calls",
                       "lockstep": "one ilqr_solve of the whole batch (every problem to the reference's own termination)",
                       "mpc": "one receding-horizon step: plant step with the first action, shift, roll out, warm-started solve! of every "
                              "problem (ilqr_mpc_run: problems advance at their own pace)"}[mode],
              "batch_per_gpu": B, "T": T, "options": "reference defaults (src/options.jl)",
              "l2": "per-tick working set (>= 1.5 GB) exceeds the 126 MB L2; no flush"}
    if mode == "stream":
        config["slots"] = args.slots

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(cfg, sample, max(1, min(K, 50) if mode != "mpc" else min(K, 20)), 1)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "iterations_per_unit": r["iterations_per_unit"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ native arm
    import torch
    import torch.distributed as dist

    from ilqr_b200 import build, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model, _, _, _ = config_inputs(cfg, 1, 0)
    n, m, p = model.n, model.m, model.p
    lib = build.model_library(model)
    mk = lambda batch, cap: capi.Handle(lib, T, n, m, p, model.cs, model.ct, batch, device=local_rank, history_cap=cap)
    h = mk(B, 8)                                       # lock-step handle: one batch per ilqr_solve
    hs = mk(args.slots, 1) if mode == "stream" else h  # the streamed job's slots
    stream = torch.cuda.Stream(device=dev)
    h.set_stream(stream.cuda_stream)
    if hs is not h:
        hs.set_stream(stream.cuda_stream)

    # synthetic job: K steps x B problems, every (rank, step) its own seed; nominal states by open-loop rollout
    nsteps_data = {"stream": max(K, args.warmup), "lockstep": min(max(K, args.warmup), 4), "mpc": 1}[mode]
    xs, us, ws = [], [], []
    for step in range(nsteps_data):
        _, x1, ubar, w = config_inputs(cfg, B, job_seed(cfg, rank, step))
        if w is not None and not ws:
            h.set_parameters(w); ws.append(w)  # the same parameters for every step of this rank (see config_inputs)
        xs.append(h.rollout(x1, ubar)); us.append(ubar)
    hx = torch.from_numpy(np.concatenate(xs)).pin_memory()      # [steps*B][T][n]  pinned host (e2e)
    hu = torch.from_numpy(np.concatenate(us)).pin_memory()
    hw = torch.from_numpy(np.concatenate(ws)).pin_memory() if ws else None
    dx, du = hx.to(dev), hu.to(dev)                              # device-resident copies (value)
    dw = hw.to(dev) if hw is not None else None
    NB = K * B if mode != "mpc" else B
    NO = max(nsteps_data * B, NB)
    ox = torch.empty((NO, T, n), dtype=torch.float64, device=dev)
    ou = torch.empty((NO, T - 1, m), dtype=torch.float64, device=dev)
    oit = torch.zeros(NO, dtype=torch.int32, device=dev)
    ost = torch.zeros(NO, dtype=torch.uint8, device=dev)
    oJ = torch.zeros(NO, dtype=torch.float64, device=dev)
    omv = torch.zeros(NO, dtype=torch.float64, device=dev)
    if mode == "stream":
        out_hx = torch.empty((NB, T, n), dtype=torch.float64).pin_memory()
        out_hu = torch.empty((NB, T - 1, m), dtype=torch.float64).pin_memory()
    elif mode == "lockstep":  # the batch's trajectories come back into pinned host buffers as well
        out_hx = torch.empty((B, T, n), dtype=torch.float64).pin_memory()
        out_hu = torch.empty((B, T - 1, m), dtype=torch.float64).pin_memory()

    # the final gather through the C ABI: trajectories + solver scalars, preallocated
    gather_ms_holder = [0.0]
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        hs.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
        g_x = torch.empty((world, NB, T, n), dtype=torch.float64, device=dev)
        g_u = torch.empty((world, NB, T - 1, m), dtype=torch.float64, device=dev)
        g_sc = torch.empty((world, NB, 4), dtype=torch.float64, device=dev)
        sc = torch.empty((NB, 4), dtype=torch.float64, device=dev)
        ge0, ge1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gather():
        if world > 1:
            with torch.cuda.stream(stream):
                ge0.record(stream)
                torch.stack([oit[:NB].double(), ost[:NB].double(), oJ[:NB], omv[:NB]], dim=1, out=sc)
                hs.gather(ox.data_ptr(), g_x.data_ptr(), NB * T * n * 8)
                hs.gather(ou.data_ptr(), g_u.data_ptr(), NB * (T - 1) * m * 8)
                hs.gather(sc.data_ptr(), g_sc.data_ptr(), NB * 4 * 8)
                ge1.record(stream)

    tot_it = torch.zeros(B, dtype=torch.int32, device=dev)

    def job_resident(nsteps=K):
        """the engine's production mode, inputs resident in HBM"""
        if mode == "stream":
            hs.solve_stream(nsteps * B, dx.data_ptr(), du.data_ptr(), dw.data_ptr() if dw is not None else 0, ox.data_ptr(), ou.data_ptr(),
                            oit.data_ptr(), ost.data_ptr(), oJ.data_ptr(), omv.data_ptr(), 0, 0)
        elif mode == "lockstep":
            for step in range(nsteps):
                s0 = (step % nsteps_data) * B
                h.initialize_controls_device(du[s0:s0 + B].data_ptr())
                h.initialize_states_device(dx[s0:s0 + B].data_ptr())
                h.solve()
                h.get_trajectory_device(ox[step * B:(step + 1) * B].data_ptr(), ou[step * B:(step + 1) * B].data_ptr())
        else:  # mpc: nsteps receding-horizon re-solves of every problem, closed loop on the device
            h.mpc_run(nsteps, 0, 0, tot_it.data_ptr())
            h.get_trajectory_device(ox.data_ptr(), ou.data_ptr())
        if nsteps == K:
            gather()

    def job_e2e(nsteps=K):
        """the same work through the host-buffer entry points"""
        if mode == "stream":
            return hs.solve_stream_host(hx.numpy()[:NB], hu.numpy()[:NB], hw.numpy()[:NB] if hw is not None else None,
                                        out_x=out_hx.numpy(), out_u=out_hu.numpy())
        if mode == "lockstep":
            for step in range(nsteps):
                s0 = (step % nsteps_data) * B
                h.initialize_controls(hu.numpy()[s0:s0 + B]); h.initialize_states(hx.numpy()[s0:s0 + B])
                h.solve()
                x_, u_ = h.get_trajectory(out_x=out_hx.numpy(), out_u=out_hu.numpy())
            return None
        for step in range(nsteps):  # ilqr_mpc_step: applied action and next plant state come back to the host every step
            h.mpc_step()
        return None

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
        own = e0.elapsed_time(e1)
        ms = torch.tensor([own], dtype=torch.float64, device=dev)
        per_rank = [own]
        if world > 1:
            allms = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allms, ms)
            per_rank = [float(t.item()) for t in allms]
        return max(per_rank), per_rank

    def mpc_reset():
        h.initialize_controls_device(du[:B].data_ptr()); h.initialize_states_device(dx[:B].data_ptr())
        h.solve()   # the initial (cold) solve is not part of the timed closed loop

    # warm-up: W steps' worth through every timed path (graph capture, allocator, clocks), the gather included
    if mode == "mpc":
        mpc_reset()
    job_resident(args.warmup)
    gather()
    if mode == "stream":
        steps_lockstep()
    K_e2e = K if mode != "mpc" else min(K, 50)
    if mode == "mpc":
        mpc_reset()
    job_e2e(min(K_e2e, args.warmup) if mode != "stream" else K)

    clocks = ClockSampler(local_rank) if rank == 0 else None
    if mode == "mpc":
        mpc_reset()
    launches0 = int(hs.get_counters()["launches"])
    ms_value, per_rank_ms = timed(job_resident)          # headline: inputs resident in HBM
    launches_timed = int(hs.get_counters()["launches"]) - launches0
    gather_ms = float(ge0.elapsed_time(ge1)) if world > 1 else 0.0
    iters = {"stream": lambda: oit[:NB].cpu().numpy().copy(), "mpc": lambda: tot_it.cpu().numpy().copy(),
             "lockstep": lambda: h.get_stats()["iterations"].copy()}[mode]()
    viol = omv[:NB].cpu().numpy().copy() if mode == "stream" else h.get_stats()["max_violation"]
    got_x = ox[:sample].cpu().numpy().copy(); got_u = ou[:sample].cpu().numpy().copy()
    got_it = oit[:sample].cpu().numpy().copy() if mode == "stream" else h.get_stats()["iterations"][:sample].copy()
    got_J = oJ[:sample].cpu().numpy().copy() if mode == "stream" else h.get_stats()["objective"][:sample].copy()
    if mode == "mpc":
        mpc_reset()
    ms_e2e, _ = timed(lambda: job_e2e(K_e2e))            # same job through the host-buffer C ABI calls
    ms_lock = timed(steps_lockstep)[0] if mode == "stream" else None
    if mode == "mpc":
        mpc_reset()
    hs.set_profiling(True)                               # the job again with CUDA events around every kernel
    ms_prof, _ = timed(lambda: job_resident(K if mode != "mpc" else min(K, 20)))
    counters = hs.get_counters()                         # (set_profiling(True) zeroes the counters: they cover this pass)
    hs.set_profiling(False)
    prof_steps = K if mode != "mpc" else min(K, 20)
    clock_info = clocks.stop() if clocks else None

    units = NB if mode != "mpc" else B * K
    total = units * world
    value = total / (ms_value * 1e-3)
    e2e_units = (NB if mode != "mpc" else B * K_e2e) * world
    e2e_value = e2e_units / (ms_e2e * 1e-3)
    if mode == "mpc":
        h2d, d2h = 0, K_e2e * B * (m + n) * 8
    else:
        h2d = (NB * T * n + NB * (T - 1) * m) * 8
        d2h = h2d + (NB * (4 + 1 + 8 + 8 + 8 + 4) if mode == "stream" else 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the profiled pass)
    names = ["forward", "linearize", "backward"]
    fused = mode != "lockstep" or model.n * model.n <= 100
    ab = algorithmic_bytes(model, T, fused=fused)
    kms = [float(v) for v in counters["kernel_ms"]]
    kl = [int(v) for v in counters["kernel_launches"]]
    pt = int(counters["problem_ticks"])
    peak, peak_src = hbm_peak()
    kernels = {}
    for i, nm in enumerate(names):
        gbs = ab[nm] * pt / (kms[i] * 1e-3) / 1e9 if kms[i] > 0 else 0.0
        kernels[nm] = {"ms_total": kms[i], "launches": kl[i], "us_per_launch": 1e3 * kms[i] / max(kl[i], 1),
                       "share_of_step": kms[i] / ms_prof, "algorithmic_bytes_per_problem_tick": ab[nm],
                       "ns_per_problem_tick": 1e6 * kms[i] / max(pt, 1), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        if mode == "lockstep":  # launches in which the batch sits a kernel out (first forward, last backward of a solve) take ~0 ms:
            kernels[nm]["ms_per_launch_all_problems_working"] = 1e6 * kms[i] / max(pt, 1) * B / 1e6  # ... so this, not us_per_launch
        if nm == "forward" and kms[i] > 0:  # the same kernel with the expected-decrease sweep's inputs counted as algorithmic bytes
            g2 = (ab[nm] + ab["forward_dgp_inputs"]) * pt / (kms[i] * 1e-3) / 1e9
            kernels[nm].update({"algorithmic_bytes_incl_expected_decrease_inputs": ab[nm] + ab["forward_dgp_inputs"],
                                "achieved_gbs_incl_expected_decrease_inputs": g2, "frac_incl_expected_decrease_inputs": g2 / peak})
    dom = max(names, key=lambda nm: kernels[nm]["ms_total"])
    if args.config == "c4":
        dom = "backward"  # north_star: the Riccati kernel against the FP64 roofline (the other kernels are listed beside it)
    ticks = int(counters["ticks"])
    kname = {"forward": "k_forward", "linearize": "k_linearize",
             "backward": "k_linback (fused gradients!+backward_pass!)" if fused else "k_backward (Riccati, CTA per problem)"}
    traffic, traffic_n, traffic_file = None, None, None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/summarize.py)
        traffic_file = "profiles/r2_traffic_c4.json" if args.config == "c4" else "profiles/r2_traffic.json"
        with open(os.path.join(ROOT, traffic_file)) as f:
            tj = json.load(f)
        if args.config == tj.get("config", "c2"):
            base = {"forward": "k_forward", "backward": "k_linback" if fused else "k_backward", "linearize": "k_linearize"}[dom]
            hits = [v for k, v in tj["dram_bytes_per_launch"].items() if k.startswith(base)]  # k_forward_tma, k_linback_tp, ...
            traffic = hits[0] if hits else None
            traffic_n = int(tj.get("problems_per_launch", 4096))
    except Exception:
        pass
    if fused or dom != "backward":
        roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic,
                    "traffic_note": None if traffic is None else
                        f"bytes per launch with all {traffic_n} problems iterating (ncu --set full, {traffic_file}); "
                        f"algorithmic bytes per such launch = algorithmic_bytes_per_problem_tick x {traffic_n}",
                    "peak_source": peak_src}
    else:  # wide-model Riccati kernel: FP64-bound (19.6 flop/B)
        fpeak, fsrc = fp64_peak()
        tf = riccati_flops(model, T) * pt / (kms[2] * 1e-3) / 1e12
        roofline = {"bound": "fp64", "kernel": kname[dom], "achieved": tf, "peak": fpeak, "unit": "TFLOP/s", "frac": tf / fpeak,
                    "traffic": traffic, "peak_source": fsrc,
                    "note": "tcgen05 has no FP64 kind; the kernel's contractions are DMMA (mma.sync m8n8k4) tiles: 37.1 TFLOP/s measured on this "
                            "GPU (DFMA 33.7); flops = the algorithm's (src/backward_pass.jl:52-84), not the tiles issued"}
    roofline.update({"how": "algorithmic bytes (flops) per problem-tick (SURVEY 8d) x problem-ticks / sum of that kernel's CUDA-event "
                            "durations on the solve stream, over a second pass of the same job with events around every kernel "
                            "(ms_per_step_with_kernel_events); the headline pass runs the same kernels from CUDA graphs.  The fused "
                            "k_linback (gradients! + backward_pass!) is listed under 'backward' with SURVEY 8d's fused byte count; models "
                            "with constant stage Hessians keep ONE accumulator per problem (HACC), which removes 2H of those bytes per "
                            "step from the actual traffic, so `frac` can exceed what the DRAM pipe itself moved",
                     "ms_per_step_with_kernel_events": ms_prof / prof_steps, "kernels": kernels})

    # ---- CPU baseline on the same inputs, and parity of the GPU results against it
    cpu, parity = None, None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(cfg, sample, min(K, 3), 1, keep_first=True)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "iterations_per_unit")}
        f = r["first"]
        if mode == "lockstep" and f is not None:
            # a FRESH handle for the sample (an unconstrained solver's iteration counter is cumulative over the solves of a
            # handle, src/solve.jl:137-139, and the timed handle has solved K batches by now)
            hp = mk(sample, 1)
            _, x1s, ubs, wsmp = sample_inputs(cfg, sample, 0)
            if wsmp is not None:
                hp.set_parameters(wsmp)
            xbs = hp.rollout(x1s, ubs)
            hp.initialize_controls(ubs); hp.initialize_states(xbs); hp.solve()
            got_x, got_u = hp.get_trajectory()
            stp = hp.get_stats()
            got_it, got_J = stp["iterations"], stp["objective"]
            hp.close()
        if mode != "mpc" and f is not None:
            it_eq = int(np.sum(got_it == f["iterations"]))
            parity = {"checked": int(sample), "against": "oracle/ilqr_oracle.c on the same inputs (rank 0, step 0)",
                      "iterations_equal": it_eq, "x_max_abs": float(np.max(np.abs(got_x - f["x"]))),
                      "u_max_abs": float(np.max(np.abs(got_u - f["u"]))),
                      "objective_max_rel": float(np.max(np.abs(got_J - f["objective"]) / np.maximum(np.abs(f["objective"]), 1e-300))),
                      "bitwise_equal": bool(np.array_equal(got_x, f["x"]) and np.array_equal(got_u, f["u"])),
                      "north_star": "identical iteration counts, trajectories to 1e-7",
                      "ok": bool(it_eq == sample and np.max(np.abs(got_x - f["x"])) <= 1e-7 and np.max(np.abs(got_u - f["u"])) <= 1e-7)}
        elif mode == "mpc":
            # closed-loop parity on a separate small handle: applied actions / plant states of the first steps
            hp = mk(sample, 1)
            _, x1, ubar, _ = sample_inputs(cfg, sample, 0)
            xb = hp.rollout(x1, ubar)
            hp.initialize_controls(ubar); hp.initialize_states(xb); hp.solve()
            nchk = len(f["applied_u"])
            au = torch.zeros((nchk, sample, m), dtype=torch.float64, device=dev)
            xn = torch.zeros((nchk, sample, n), dtype=torch.float64, device=dev)
            hp.mpc_run(nchk, au.data_ptr(), xn.data_ptr(), 0)
            torch.cuda.synchronize(dev)
            da = max(float(np.max(np.abs(au[s].cpu().numpy() - f["applied_u"][s]))) for s in range(nchk))
            dxn = max(float(np.max(np.abs(xn[s].cpu().numpy() - f["x_next"][s]))) for s in range(nchk))
            parity = {"checked": int(sample), "steps": nchk, "against": "oracle closed loop on the same inputs (seed 0)",
                      "applied_action_max_abs": da, "plant_state_max_abs": dxn, "ok": bool(da <= 1e-7 and dxn <= 1e-7),
                      "bitwise_equal": bool(da == 0.0 and dxn == 0.0)}
            hp.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "ms_per_iteration": ms_prof / max(ticks, 1),
            "ticks_per_step": ticks / prof_steps,
            "slot_fill": pt / max(ticks * (args.slots if mode == "stream" else B), 1),
            "compactions": int(counters.get("compactions", 0)),
            "per_rank_ms": per_rank_ms, "gather_ms": gather_ms,
            "iterations_per_problem": {"mean": float(np.mean(iters)) / (K if mode == "mpc" else 1),
                                       "max": float(np.max(iters)) / (K if mode == "mpc" else 1),
                                       "what": "per re-solve, averaged over the K steps of each problem" if mode == "mpc" else "per solve"},
            "converged_frac": float((np.asarray(viol) <= 5e-3).mean()),
            "problems_per_step_per_gpu": B,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d // max(K_e2e if mode == "mpc" else K, 1)),
                    "d2h_bytes_per_step": int(d2h // max(K_e2e if mode == "mpc" else K, 1)), "ms_per_step": ms_e2e / (K_e2e if mode == "mpc" else K),
                    "steps": K_e2e if mode == "mpc" else K,
                    "call": {"stream": "ilqr_solve_stream_host (pinned host buffers in, host buffers out)",
                             "lockstep": "ilqr_initialize_controls/states + ilqr_solve + ilqr_get_trajectory (host buffers)",
                             "mpc": "ilqr_mpc_step per step (applied action and next plant state returned to the host)"}[mode]},
            "gpu_launches": launches_timed,
            "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
    if ms_lock is not None:
        line["lockstep_batches"] = {"value": total / (ms_lock * 1e-3), "unit": UNIT, "ms_per_step": ms_lock / args.steps,
                                    "what": "K separate ilqr_solve calls of one batch each: every batch waits for its slowest problem"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
