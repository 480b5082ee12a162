"""The N>1 host path on CPU: two gloo ranks shard one array contiguously, evaluate their shard
(with the oracle standing in for the device), and reduce the timing scalar with MAX -- the same
bookkeeping bench.py does under torchrun with NCCL.  No data-path collective exists."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, 'tests', 'golden'))
    import golden_io
    import numpy as np
    import torch.distributed as dist
    from adaptive_interpolation_b200 import sharding, tables
    from oracle import oracle_np
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = golden_io.load_golden("approx__64o7-p1.19209289551e-06")
    f = tables.flatten(g)
    n = 100003
    x = np.random.default_rng(7).uniform(0, 20, n)
    lo, hi = sharding.shard_of(n, rank, world)
    y_local = oracle_np.evaluate(f.breaks, f.coef, f.basis, f.order, x[lo:hi])
    t = sharding.max_over_ranks(1.0 + rank, dist)
    pts = sharding.sum_over_ranks(hi - lo, dist)
    assert t == float(world), t
    assert pts == n, pts
    y_full = oracle_np.evaluate(f.breaks, f.coef, f.basis, f.order, x)
    assert np.array_equal(y_local, y_full[lo:hi])
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok", lo, hi)
""") % (ROOT, ROOT)


@pytest.mark.timeout(180)
def test_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=170)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [0, 1, 511, 512, 100003, 1 << 20, 1 << 30])
def test_shard_bounds_cover_the_array(n, world):
    """bench.py's strong-scaled configs (BASELINE configs[2] / [4]: 2^30 GLOBAL points over the N
    GPUs) and generate.run_multi split one array with sharding.shard_bounds: contiguous, disjoint,
    covering, interior boundaries on multiples of 512 elements (every shard keeps the 16-byte vector
    path), and equal shards whenever the size allows."""
    from adaptive_interpolation_b200 import sharding
    b = sharding.shard_bounds(n, world)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
    for (lo, hi), (lo2, _) in zip(b, b[1:]):
        assert lo <= hi == lo2
    for lo, hi in b[:-1]:
        assert hi % 512 == 0 or n < 512 * world
    if n == 1 << 30 and world in (1, 2, 4, 8):
        assert all(hi - lo == n // world for lo, hi in b)
    assert [sharding.shard_of(n, r, world) for r in range(world)] == b
