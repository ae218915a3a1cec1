#!/usr/bin/env python
"""tools/prof_target.py WHAT -- a short program for ncu: a few launches of one kernel family.
WHAT: playout | greedy | greedy16 | learn | table"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subproc_b200 import ops, parameter, value_table

what = sys.argv[1]
dev = torch.device("cuda:0")
w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(dev)
if what == "playout":
    po = ops.playout(1 << 20, seed=1, gid0=0, device=dev)
    for i in range(3):
        ops.playout(1 << 20, seed=1, gid0=(i + 1) << 20, device=dev, out=po)
elif what.startswith("greedy"):
    n = 1 << (16 if what == "greedy16" else 19)
    po = ops.playout(n, seed=2, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
    for i in range(2):
        ops.playout(n, seed=2, gid0=(i + 1) * n, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w, out=po)
elif what == "learn":
    po = ops.playout(1 << 16, seed=3, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
    for i in range(3):
        ops.learn_accumulate(po)
elif what == "table":
    po = ops.playout(1 << 16, seed=2, gid0=0, device=dev)
    vt = value_table.ValueTable(device=dev)
    vt.update_from_playout(po)
    vt.update_from_playout(po)
torch.cuda.synchronize()
print("done", what)
