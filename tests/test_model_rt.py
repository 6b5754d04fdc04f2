"""include/ilqr_model_rt.h: the portable sin/cos used by generated model code."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    d = tempfile.mkdtemp()
    src = os.path.join(d, "rt.c")
    with open(src, "w") as f:
        f.write('#include "ilqr_model_rt.h"\n'
                'void sc(const double* x, double* s, double* c, int n){for(int i=0;i<n;i++) ilqr_sincos(x[i],s+i,c+i);}\n')
    lib = os.path.join(d, "rt.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", f"-I{ROOT}/include", src, "-o", lib, "-lm"])
    return ctypes.CDLL(lib)


def test_sincos_accuracy_against_mpmath():
    import mpmath as mp
    lib = _build()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-10, 10, 3000), rng.uniform(-1e5, 1e5, 1500), rng.normal(0, 1e-3, 300),
                         np.arange(-40, 41) * np.pi / 2, np.arange(-40, 41) * np.pi / 4, [0.0, -0.0, 1e-300, 105614.9]])
    n = len(xs)
    s, c = np.zeros(n), np.zeros(n)
    P = ctypes.POINTER(ctypes.c_double)
    lib.sc(xs.ctypes.data_as(P), s.ctypes.data_as(P), c.ctypes.data_as(P), n)
    mp.mp.prec = 200
    worst = 0.0
    for i in range(n):
        x = mp.mpf(float(xs[i]))
        for got, true in ((s[i], mp.sin(x)), (c[i], mp.cos(x))):
            ft = float(true)
            ulp = np.spacing(abs(ft)) if ft != 0 else 5e-324
            worst = max(worst, float(abs((mp.mpf(float(got)) - true) / mp.mpf(float(ulp)))))
    assert worst < 2.0, worst  # CUDA documents 2 ulp for its own sin/cos


def test_sincos_special_values():
    lib = _build()
    xs = np.array([0.0, -0.0, np.inf, np.nan, 1e9])
    s, c = np.zeros(5), np.zeros(5)
    P = ctypes.POINTER(ctypes.c_double)
    lib.sc(xs.ctypes.data_as(P), s.ctypes.data_as(P), c.ctypes.data_as(P), 5)
    assert s[0] == 0.0 and c[0] == 1.0
    assert s[1] == 0.0  # the sign of zero is not preserved (documented in DESIGN.md)
    assert np.isnan(s[2]) and np.isnan(s[3])
    assert abs(s[4] - np.sin(1e9)) < 1e-15  # platform fallback outside the reduced range
