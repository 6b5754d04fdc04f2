"""Build recipe for the C oracle (test infrastructure).

    python oracle/build_oracle.py            # builds the oracle for every fixture model

Compiles oracle/ilqr_oracle.c against a GENERATED model header (the same text the CUDA
engine is compiled against) with implicit FMA contraction off -- the arithmetic contract
of DESIGN.md.  Output: oracle/_build/<model>_<hash>/liboracle.so (git-ignored)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CFLAGS = ["-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-Wno-unused-function"]


def build_oracle(name: str, header_text: str, digest: str, force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build", f"{name}_{digest}")
    os.makedirs(out_dir, exist_ok=True)
    hdr = os.path.join(out_dir, "model.h")
    lib = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(HERE, "ilqr_oracle.c")
    deps = [src, os.path.join(ROOT, "include", "ilqr_model_rt.h"), os.path.join(ROOT, "include", "ilqr_cuda.h")]
    if not os.path.exists(hdr) or open(hdr).read() != header_text:
        with open(hdr, "w") as f:
            f.write(header_text)
    cmd = ["gcc", *CFLAGS, f"-I{os.path.join(ROOT, 'include')}", "-include", hdr, src, "-o", lib, "-lm"]
    # content stamp (command line + source bytes), not modification times: those do not survive the copy to the GPU box
    h = hashlib.sha256(" ".join(cmd).replace(ROOT, "$ROOT").encode())  # the tree may sit elsewhere on the GPU box
    for dep in deps + [hdr]:
        with open(dep, "rb") as f:
            h.update(hashlib.sha256(f.read()).digest())
    stamp = h.hexdigest()
    try:
        with open(lib + ".stamp") as f:
            fresh = os.path.exists(lib) and f.read().strip() == stamp
    except OSError:
        fresh = False
    if fresh and not force:
        return lib
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{' '.join(cmd)}\n{r.stderr}")
    with open(lib + ".stamp", "w") as f:
        f.write(stamp)
    return lib


def build_for_model(model, force: bool = False) -> str:
    return build_oracle(model.name, model.header, model.hash, force)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    import ilqr_b200  # noqa: F401
    from ilqr_b200 import problems

    for ctor in (problems.particle, problems.pendulum, problems.car, problems.acrobot):
        print(build_for_model(ctor(), force="--force" in sys.argv))
