"""small run of every kernel for compute-sanitizer (memcheck / racecheck), one tool per gpurun call"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from subproc_b200 import ops, value_table, learner, parameter
dev = "cuda:0"
w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(dev)
po = ops.playout(4096 + 5, seed=1, gid0=0, device=dev)
pg = ops.playout(2048 + 3, seed=2, gid0=0, device=dev, policy=ops.POLICY_GREEDY, random_plies=4, n_rand_black=2, n_rand_white=1, weights=w)
b, wh = po.black[20].contiguous(), po.white[20].contiguous()
n = b.numel()
side = torch.ones(n, dtype=torch.uint8, device=dev)
ops.legal(b, wh); ops.flips(b, wh, po.move[20].contiguous()); ops.counts(b, wh); ops.features(b, wh, side); ops.evaluate(b, wh, side, w)
ops.step(b.clone(), wh.clone(), side.clone(), torch.zeros(n, dtype=torch.int32, device=dev), po.move[20].contiguous())
ops.deserialize_boards(ops.serialize_boards(b, wh))
ops.learn_accumulate(pg)
vt = value_table.ValueTable(device=dev); vt.update_from_playout(pg); vt.update_from_playout(po)
print("perft7", ops.perft(7, device=dev))
torch.cuda.synchronize()
print("sanitize run ok", po.total_positions(), pg.total_positions(), len(vt))
