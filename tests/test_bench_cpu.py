"""bench.py's accounting (no GPU): the algorithmic byte / flop counts are SURVEY.md 8d's, the roofline denominators and the
committed DRAM-traffic files resolve for the kernels the engine actually launches."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import ilqr_b200  # noqa: E402,F401
from ilqr_b200 import problems  # noqa: E402


def test_algorithmic_counts_are_the_surveys():
    ac = problems.acrobot()
    ab = bench.algorithmic_bytes(ac, 101, fused=True)
    # SURVEY 8d, acrobot T = 101: forward 120 B per step (x, u, K, k read; x, u written ...), fused linearise + Riccati
    assert ab["forward"] == 12064 and ab["backward"] == 45984 and ab["linearize"] == 0
    assert ab["forward"] + ab["forward_dgp_inputs"] == 32064  # + fx, fu, Lx, Lu of every step (expected-decrease sweep)
    lq = problems.lq_tracking(8, 2)
    n, m, T = 8, 2, 16
    assert bench.riccati_flops(lq, T) == (4 * n**3 + 10 * n * n * m + 6 * n * m * m + m**3 / 3) * (T - 1)


def test_roofline_denominators_and_traffic_files():
    peak, src = bench.fp64_peak()
    with open(os.path.join(ROOT, "profiles", "r2_fp64_peak.json")) as f:
        pk = json.load(f)
    assert peak == pk["dmma_m8n8k4_tflops"] > pk["dfma_tflops"] > 30 and "measured" in src  # the tensor pipe the Riccati tiles run on
    hbm, _ = bench.hbm_peak()
    assert 5000 < hbm < 8000
    for name, config, bases in (("r2_traffic.json", "c2", ("k_forward", "k_linback")), ("r2_traffic_c4.json", "c4", ("k_backward", "k_forward"))):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            tj = json.load(f)
        assert tj["config"] == config and tj["problems_per_launch"] > 0
        for base in bases:  # bench.py looks the dominant kernel up by prefix (k_forward_tma, k_linback_tp, ...)
            hits = [v for k, v in tj["dram_bytes_per_launch"].items() if k.startswith(base)]
            assert hits and hits[0] > 1e8, (name, base)


def test_configs_name_the_baseline_workloads():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert set(bench.CONFIGS) == {"c2", "c3", "c4", "c5"}
    assert bench.CONFIGS["c2"]["batch"] == 4096 and bench.CONFIGS["c2"]["T"] == 101
    assert bench.CONFIGS["c4"]["T"] == 256 and bench.CONFIGS["c4"]["batch"] == 1024
    assert "metric" in base and bench.CONFIGS["c2"]["metric"]
    for c in bench.CONFIGS.values():  # distinct inputs per (rank, step)
        assert bench.job_seed(c, 0, 0) != bench.job_seed(c, 1, 0) or c["mode"] != "stream"
