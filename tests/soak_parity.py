#!/usr/bin/env python
"""tests/soak_parity.py -- long differential run: GPU playouts vs the oracle, bit for bit, on far more
games than the test suite plays (random, go_for substitution, greedy, mixed engines, custom starts)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from subproc_b200 import ops
from oracle import lib as orc

dev = "cuda:0"
w = torch.from_numpy(orc.DEFAULT_WEIGHTS.astype(np.float32)).to(dev)
rng = np.random.RandomState(0)
other = np.concatenate([rng.randint(-60, 100, size=(4, 9)).astype(np.float64), np.zeros((4, 1))], axis=1)
wo = torch.from_numpy(other.astype(np.float32)).to(dev)
cases = [
    ("random", 200000, dict(), dict()),
    ("greedy R=10", 20000, dict(policy=1, random_plies=10), dict(weights=w)),
    ("greedy + go_for substitution 10/2", 10000, dict(policy=1, n_rand_black=10, n_rand_white=2), dict(weights=w)),
    ("greedy(other) vs random", 10000, dict(policy=1, policy_white=0, random_plies=2, weights=other), dict(weights=wo)),
    ("greedy(default) vs greedy(other)", 10000, dict(policy=1, policy_white=1, random_plies=4, weights_white=other),
     dict(weights=w, weights_white=wo)),
]
total_games = total_plies = 0
t0 = time.time()
for name, n, okw, gkw in cases:
    seed = 1000 + len(name)
    ref = orc.playout(seed, 0, n, **okw)
    gk = {k: v for k, v in okw.items() if k not in ("weights", "weights_white")}
    po = ops.playout(n, seed=seed, gid0=0, device=dev, **gk, **gkw)
    npl = po.nplies.cpu().numpy()
    assert np.array_equal(npl, ref['nplies']), name
    assert np.array_equal(ops.bits_numpy(po.final_black), ref['final_black']), name
    assert np.array_equal(ops.bits_numpy(po.final_white), ref['final_white']), name
    t_idx = np.arange(po.t_max + 1)[:, None]
    vp = t_idx <= npl[None, :]
    vm = t_idx[:-1] < npl[None, :]
    assert np.array_equal(ops.bits_numpy(po.black)[vp], ref['black'][vp]), name
    assert np.array_equal(ops.bits_numpy(po.white)[vp], ref['white'][vp]), name
    assert np.array_equal(po.move.cpu().numpy()[vm], ref['move'][vm]), name
    total_games += n
    total_plies += int(npl.sum())
    print("%-40s %7d games %9d plies  passes %6d  bit-exact (every position and move)"
          % (name, n, int(npl.sum()), int((ref['move'][vm] == 64).sum())))
print("soak ok: %d games, %d plies compared in %.0f s" % (total_games, total_plies, time.time() - t0))
