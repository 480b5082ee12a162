"""bench.py's synthetic point generators (CPU tensors here; the bench fills device tensors with the
same code).  BASELINE configs[3] asks for points CLUSTERED at the breakpoints: the device generator
must be the distribution tests/golden/make_golden.py:make_points(clustered=True) draws the parity
fixtures from."""
import numpy as np
import pytest
import torch

import bench
from adaptive_interpolation_b200 import tables
from conftest import golden


@pytest.mark.parametrize("name", ["gen__cfg4_f1_leg_o7_f32", "gen__cfg4_f1_leg_o7_f64"])
def test_clustered_points_follow_the_config_4_recipe(name):
    flat = tables.flatten(golden(name))
    n = 1 << 18
    x = torch.empty(n, dtype=torch.float32 if flat.dtype == np.float32 else torch.float64)
    gen = torch.Generator(device="cpu")
    gen.manual_seed(5)
    bench.fill_points(torch, x, flat, "clustered", gen)
    xs = x.numpy().astype(np.float64)
    b = flat.breaks.astype(np.float64)
    lb, ub = b[0], b[-1]
    # random part: r >= ub -> nextafter(ub, -inf); only the planted ub and its upper neighbour reach ub
    assert np.all(np.isfinite(xs)) and np.sum(xs >= ub) <= 2 and xs.max() <= np.nextafter(flat.dtype.type(ub), np.inf)
    # distance to the nearest breakpoint: half of the points are +-|N(0,1)| * 2^-j, j in [10, 40]
    pos = np.searchsorted(b, xs)
    d = np.minimum(np.abs(xs - b[np.clip(pos - 1, 0, len(b) - 1)]), np.abs(xs - b[np.clip(pos, 0, len(b) - 1)]))
    near = np.mean(d < 4 * 2.0 ** -10)
    assert 0.45 < near < 0.75, near                                      # ~1/2 clustered + the uniform ones that fall close
    very_near = np.mean(d < 2.0 ** -30)
    assert very_near > 0.08, very_near                                   # j reaches 40: a third of the clustered half
    # the uniform half covers the whole domain
    hist, _ = np.histogram(xs, bins=10, range=(lb, ub))
    assert hist.min() > 0.02 * n
    # every breakpoint and both neighbours are planted (collisions of the random positions aside)
    dt = flat.dtype.type
    want = np.concatenate([flat.breaks, np.nextafter(flat.breaks, dt(-np.inf)), np.nextafter(flat.breaks, dt(np.inf))])
    have = set(x.numpy().tolist())
    assert np.mean([w in have for w in want.tolist()]) > 0.95


def test_uniform_and_sorted_points():
    flat = tables.flatten(golden("approx__32o6-p1.19209289551e-06"))
    gen = torch.Generator(device="cpu")
    gen.manual_seed(1)
    x = torch.empty(1 << 16, dtype=torch.float32)
    bench.fill_points(torch, x, flat, "uniform", gen)
    assert float(x.min()) >= 0.0 and float(x.max()) < 20.0 and abs(float(x.mean()) - 10.0) < 0.2
    parts = []
    for rank in range(2):                                                # two ranks: consecutive halves of one linspace
        x = torch.empty(1 << 12, dtype=torch.float32)
        bench.fill_points(torch, x, flat, "sorted", gen, rank=rank, world=2)
        parts.append(x.numpy())
    both = np.concatenate(parts)
    assert np.all(np.diff(both) >= 0) and both[0] == 0.0 and both[-1] < 20.0 and parts[1][0] >= 10.0
    with pytest.raises(ValueError):
        bench.fill_points(torch, x, flat, "zigzag", gen)
