#!/usr/bin/env python
"""tests/run_all_configs.py -- the five BASELINE.json configs, each checked and timed on ONE GPU
(configs 4/5 at their per-GPU size; tools/bench_configs.py runs them over torchrun for N GPUs).
Prints one JSON object; `parity` entries are checked against the oracle / golden vectors here."""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from subproc_b200 import ops, learner, parameter
from oracle import lib as orc

dev = "cuda:0"
out = {}


def ev_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


# config 1: one seeded random-vs-random game, the reference's board.py on the CPU vs the kernel
g = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "games.json.gz")).read())[0]
po = ops.playout(1, seed=g['seed'], gid0=g['gid'], device=dev)
n = len(g['plies'])
same = (int(po.nplies.cpu()[0]) == n and po.move[:n, 0].cpu().tolist() == [p['move'] for p in g['plies']]
        and [int(v) for v in ops.bits_numpy(po.black[:n + 1, 0].contiguous())] == [int(p['b'], 16) for p in g['positions']])
c = po.final_counts().cpu().numpy()[0]
out["config1_single_game"] = {"plies": n, "final_black_white": [int(c[0]), int(c[1])],
                              "parity": "identical to the game board.py played (tests/golden/games.json.gz[0])" if same else "MISMATCH"}

# config 2: perft 1..10 from the start position
want = [4, 12, 56, 244, 1396, 8200, 55092, 390216, 3005288, 24571284]
got = [ops.perft(d, device=dev) for d in range(1, 11)]
t0 = time.perf_counter(); ops.perft(10, device=dev); dt = time.perf_counter() - t0
out["config2_perft"] = {"depth_1_to_10": got, "parity": "bit-exact" if got == want else "MISMATCH",
                        "perft10_ms": dt * 1e3, "perft10_nodes_per_s": want[-1] / dt}

# config 3: 2^20 lock-step random playouts, full trajectories
B = 1 << 20
po = ops.playout(B, seed=1, gid0=0, device=dev)
ms = ev_ms(lambda: ops.playout(B, seed=1, gid0=0, device=dev, out=po))
pos = po.total_positions()
k = 2048
ref = orc.playout(1, 0, k)
ok = (np.array_equal(po.nplies[:k].cpu().numpy(), ref['nplies']) and
      np.array_equal(ops.bits_numpy(po.final_black[:k]), ref['final_black']))
out["config3_random_playouts"] = {"games": B, "positions": pos, "kernel_ms": ms, "positions_per_s": pos / ms * 1e3,
                                  "games_per_s": B / ms * 1e3, "parity": "%d games bit-exact vs oracle" % k if ok else "MISMATCH"}

# config 4: greedy self-play on default_value() weights, 2^19 games per GPU, first 10 plies random
G = 1 << 19
w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(dev)
pg = ops.playout(G, seed=2, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
ms = ev_ms(lambda: ops.playout(G, seed=2, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w, out=pg), reps=3)
pos = pg.total_positions()
k = 512
ref = orc.playout(2, 0, k, policy=1, random_plies=10)
ok = np.array_equal(pg.nplies[:k].cpu().numpy(), ref['nplies']) and np.array_equal(ops.bits_numpy(pg.final_black[:k]), ref['final_black'])
out["config4_greedy_selfplay"] = {"games_per_gpu": G, "positions": pos, "kernel_ms": ms, "positions_per_s": pos / ms * 1e3,
                                  "games_per_s": G / ms * 1e3, "parity": "%d games bit-exact vs oracle" % k if ok else "MISMATCH"}

# config 5: one learner iteration (self-play + statistics + refit); the all-reduce is a no-op at N = 1
L = learner.ProgressPositionMovesLearn(); L.configure({})
L.self_play_iteration(1 << 16, seed=3, iteration=0, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(1, 6):
    L.self_play_iteration(1 << 16, seed=3, iteration=it, device=dev)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
out["config5_learner_iteration"] = {"games_per_gpu": 1 << 16, "iteration_ms": dt * 1e3, "games_per_s": (1 << 16) / dt,
                                    "parameters": list(L.read_parameters()),
                                    "fit_r2": [round(f['r2'], 4) for f in L.last_fits]}
print(json.dumps(out))
