"""The N > 1 host logic on CPU: two gloo ranks shard a batch, each 'solves' its shard (here:
the C oracle stands in for the per-GPU engine -- test infrastructure), and the gather puts the
results back in batch order on every rank."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))
    import numpy as np, torch, torch.distributed as dist
    import ilqr_b200
    from ilqr_b200.distributed import shard_bounds, split_inputs, gather_shards
    from common import inputs
    from oracle.c_oracle import COracle
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    B, T = 7, 11
    model, x1, ubar = inputs('particle', B, T, seed=3)
    mine = split_inputs({'x1': x1, 'u': ubar}, world, rank)
    lo, hi = shard_bounds(B, world, rank)
    assert mine['x1'].shape[0] == hi - lo
    co = COracle(model, T, hi - lo)
    xb = co.rollout(mine['x1'], mine['u']); co.initialize_controls(mine['u']); co.initialize_states(xb); co.solve()
    x, u = co.get_trajectory(); st = co.get_stats()
    out = gather_shards({'x': torch.from_numpy(x), 'u': torch.from_numpy(u),
                         'it': torch.from_numpy(st['iterations'].astype(np.int64))}, B, dist)
    full = COracle(model, T, B)
    xbf = full.rollout(x1, ubar); full.initialize_controls(ubar); full.initialize_states(xbf); full.solve()
    xf, uf = full.get_trajectory()
    assert out['x'].shape == (B, T, 2) and np.array_equal(out['x'].numpy(), xf) and np.array_equal(out['u'].numpy(), uf)
    assert np.array_equal(out['it'].numpy(), full.get_stats()['iterations'])
    dist.barrier(); dist.destroy_process_group()
    print('rank', rank, 'ok')
""") % (ROOT, ROOT)


def test_two_rank_gloo_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2
