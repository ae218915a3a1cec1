#!/usr/bin/env python
"""tools/bench_learn.py -- the launches of one config-5 learning iteration, each timed alone (CUDA events):
greedy self-play, exact statistics (learn_kernel), statistics -> doubles, the four regressions."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subproc_b200 import ops, parameter, learner

dev = torch.device('cuda:0')
G = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(dev)


def ev(f, reps=10):
    f(); f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps


po = ops.playout(G, seed=3, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
acc = torch.zeros((4, learner.N_ACC), dtype=torch.int64, device=dev)
stats = torch.empty((4, 112), dtype=torch.float64, device=dev)
out = {"games": G, "positions": po.total_positions() + G}
out["greedy_playout_ms"] = ev(lambda: ops.playout(G, seed=3, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w, out=po))
for gpw in (32, 16, 8):
    out["greedy_playout_gpw%d_ms" % gpw] = ev(lambda: ops.playout(G, seed=3, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w, out=po, games_per_warp=gpw))
out["learn_accumulate_ms"] = ev(lambda: ops.learn_accumulate(po, acc=acc))
out["learn_positions_per_s"] = out["positions"] / out["learn_accumulate_ms"] * 1e3
out["learn_stats_ms"] = ev(lambda: ops.learn_stats(acc, out=stats))
acc.zero_(); ops.learn_accumulate(po, acc=acc); ops.learn_stats(acc, out=stats)
w2 = w.clone()
out["learn_solve_ms"] = ev(lambda: ops.learn_solve(stats, w, weights_out=w2))
out["zero_ms"] = ev(lambda: acc.zero_())
pr = ops.playout(1 << 20, seed=1, gid0=0, device=dev)
out["learn_accumulate_2^20_random_games_ms"] = ev(lambda: ops.learn_accumulate(pr, acc=acc), reps=3)
out["learn_positions_per_s_2^20"] = (pr.total_positions() + (1 << 20)) / out["learn_accumulate_2^20_random_games_ms"] * 1e3
print(json.dumps(out))
