"""Build recipes: libilqr_cuda.so (C ABI front) and per-model CUDA plug-ins.

Everything is built IN-TREE under ``iterativelqr.jl_b200/_build/`` so that the shared
objects travel with the source snapshot to the GPU box; plug-ins are cached by the hash
of their generated model header (the moral equivalent of the reference's
"#TODO: option to load/save methods", /root/reference/src/dynamics.jl:17).

nvcc flags: ``-gencode arch=compute_100a,code=sm_100a`` (B200 only), ``-fmad=false`` (the
arithmetic contract: only explicit fma fuses), ``-lineinfo`` (ncu source view).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared", "-cudart", "static"]
FRONT_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-Wall"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without the CUDA toolkit")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def _stamp(cmd, deps) -> str:
    """Content hash of everything a binary depends on: the full compiler command line and the bytes of every source.
    (Modification times do not survive the copy of the tree to the GPU box and say nothing about compiler flags.)"""
    h = hashlib.sha256(" ".join(cmd).replace(ROOT, "$ROOT").encode())  # the tree may sit elsewhere on the GPU box
    for d in deps:
        with open(d, "rb") as f:
            h.update(hashlib.sha256(f.read()).digest())
    return h.hexdigest()


def _fresh(target, stamp) -> bool:
    try:
        with open(target + ".stamp") as f:
            return os.path.exists(target) and f.read().strip() == stamp
    except OSError:
        return False


def _build(cmd, out, deps, force=False, log_path=None):
    stamp = _stamp(cmd, deps)
    if force or not _fresh(out, stamp):
        log = _run(cmd)
        with open(out + ".stamp", "w") as f:
            f.write(stamp)
        if log_path:
            with open(log_path, "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
    return out


def front_library(force: bool = False) -> str:
    """libilqr_cuda.so -- the C ABI (include/ilqr_cuda.h)."""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libilqr_cuda.so")
    src = os.path.join(CSRC, "ilqr_front.cpp")
    deps = [src, os.path.join(CSRC, "ilqr_plugin.h"), os.path.join(CSRC, "ilqr_nccl_dyn.h"), os.path.join(INCLUDE, "ilqr_cuda.h")]
    # default visibility for the extern "C" API only
    return _build(["g++", *FRONT_FLAGS, f"-I{INCLUDE}", f"-I{CSRC}", "-DILQR_BUILDING", src, "-o", out, "-ldl"], out, deps, force)


VARIANTS = {
    "": [],
    # the loop-based kernels of csrc/ilqr_large_*.cuh forced onto a small model (tests of that path)
    "large": ["-DILQR_FORCE_LARGE=1"],
    # line-search step sizes evaluated per k_forward launch (default 2; the rest of a search runs at later ticks)
    "ls1": ["-DILQR_FWD_TRIALS=1"],
    "ls4": ["-DILQR_FWD_TRIALS=4"],
    "f1": ["-DILQR_FWD_MIN_CTAS=1", "-DILQR_DG_MAX_STAGES=8", "-DILQR_PR_MAX_STAGES=6"],  # no register cap on k_forward, deep rings
    "lbtimers": ["-DILQR_LB_TIMERS=1"],
    # per-step Hessian accumulators even where one per problem would do (tests of the general path on the small fixtures)
    "nohacc": ["-DILQR_NO_HACC=1"],
    "rltimers": ["-DILQR_RL_PHASE_TIMERS=1"],  # debug: per-phase cycle counters of the wide-model Riccati kernel (printf)
    "rltimers_nodmma": ["-DILQR_RL_PHASE_TIMERS=1", "-DILQR_RL_DMMA=0"],
    "nodmma": ["-DILQR_RL_DMMA=0"],  # wide-model Riccati kernel with the register-tiled DFMA loops instead of DMMA tiles
    "tp12": ["-DILQR_TP_WARPS_PER_SM=12"],  # k_linback_tp capped at 168 registers: 12 warps per SM
    "tp10": ["-DILQR_TP_WARPS_PER_SM=10"],
    # wide models with constant dynamics Jacobians: per-(problem, step) staged blocks anyway (the general path on the LQ fixtures)
    "nojacconst": ["-DILQR_NO_JAC_CONST=1"],
    "nojacconst_nodmma": ["-DILQR_NO_JAC_CONST=1", "-DILQR_RL_DMMA=0"],
    "rltimers_nojacconst": ["-DILQR_RL_PHASE_TIMERS=1", "-DILQR_NO_JAC_CONST=1"],
    # wide-model Riccati kernel with factorisation and solves AFTER the Qxx contraction instead of under it
    "nooverlap": ["-DILQR_RL_OVERLAP=0"],
    "rl_qvec": ["-DILQR_RL_QVEC_DMMA=1"],  # A/B switches of the wide-model Riccati kernel (see ilqr_large_backward.cuh)
    "rl_nofg": ["-DILQR_RL_FG_DMMA=0"],
    "rl_helpers": ["-DILQR_RL_HELPERS=1"],        # 12 warps: factorisation and solve warps of their own, spine-first order (measured: 4 % slower)
    "rl_fwarp": ["-DILQR_RL_FWARP=1"],            # warp 0 only factorises, warp 4 takes its tiles (measured: no gain)
    "rl_cholright": ["-DILQR_RL_CHOL_RIGHT=1"],   # right-looking register Cholesky (measured: slower)
    "rl_base": ["-DILQR_RL_OVERLAP=0", "-DILQR_RL_QVEC_DMMA=0", "-DILQR_RL_FG_DMMA=0"],
    "rltimers_nooverlap": ["-DILQR_RL_PHASE_TIMERS=1", "-DILQR_RL_OVERLAP=0"],
}


def model_dir(model, variant: str = "") -> str:
    return os.path.join(BUILD, f"{model.name}_{model.hash}" + (f"_{variant}" if variant else ""))


def model_library(model, force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """Compile the CUDA plug-in for one model (cached by header hash and build variant).  ILQR_VARIANT in the
    environment replaces the default variant (experiments: run the whole test suite against another build)."""
    if not variant:
        variant = os.environ.get("ILQR_VARIANT", "")
    d = model_dir(model, variant)
    os.makedirs(d, exist_ok=True)
    hdr = os.path.join(d, "model.h")
    out = os.path.join(d, "libilqr_model.so")
    if not os.path.exists(hdr) or open(hdr).read() != model.header:
        with open(hdr, "w") as f:
            f.write(model.header)
    deps = [hdr, os.path.join(CSRC, "ilqr_engine.cu"), os.path.join(CSRC, "ilqr_kernels.cuh"),
            os.path.join(CSRC, "ilqr_large_forward.cuh"), os.path.join(CSRC, "ilqr_large_backward.cuh"),
            os.path.join(CSRC, "ilqr_plugin.h"), os.path.join(CSRC, "ilqr_nccl_dyn.h"), os.path.join(INCLUDE, "ilqr_cuda.h"),
            os.path.join(INCLUDE, "ilqr_model_rt.h")]
    # -Xptxas -v always: the register / spill report of every kernel lands in build.log next to the binary
    cmd = [_nvcc(), "-Xptxas", "-v", *NVCC_FLAGS, *VARIANTS[variant], *os.environ.get("ILQR_NVCC_EXTRA", "").split(),
           f"-I{INCLUDE}", f"-I{CSRC}", "-include", hdr, os.path.join(CSRC, "ilqr_engine.cu"), "-o", out]
    return _build(cmd, out, deps, force, os.path.join(d, "build.log"))
