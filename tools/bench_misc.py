#!/usr/bin/env python
"""tools/bench_misc.py -- timings of the remaining kernels (perft, batch rules, learner statistics,
value table, serialisation) on one GPU; prints one JSON line per kernel.  CUDA events, best of 5."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from subproc_b200 import ops, value_table

dev = torch.device("cuda:0")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


n = 1 << 20
po = ops.playout(n, seed=1, gid0=0, device=dev)
positions = po.total_positions()
t = 30
b, w = po.black[t].contiguous(), po.white[t].contiguous()
turn = torch.ones(n, dtype=torch.uint8, device=dev)
nturn = torch.zeros(n, dtype=torch.int32, device=dev)
mv = po.move[t].contiguous()
wts = ops.weights_tensor([[100, 99, -1, -1, -1, -1, 3, 8, 20], [75, 99, 2, -5, 7, 6, 4, 5, 5],
                          [25, 99, 2, -5, -7, -6, 4, 5, 5], [1, 100, 50, 30, 30, 30, 30, 30, 30]], dev)
out = []
ms = timed(lambda: ops.legal(b, w)); out.append(("othello_legal", n / ms * 1e3, "positions/s", 24 * n / ms / 1e6))
bb, ww = b.clone(), w.clone()
ms = timed(lambda: ops.step(bb.copy_(b), ww.copy_(w), turn.fill_(1), nturn, mv))
out.append(("othello_step(+2 copies)", n / ms * 1e3, "positions/s", None))
ms = timed(lambda: ops.features(b, w, turn)); out.append(("othello_features", n / ms * 1e3, "positions/s", 57 * n / ms / 1e6))
ms = timed(lambda: ops.evaluate(b, w, turn, wts)); out.append(("othello_eval", n / ms * 1e3, "positions/s", 21 * n / ms / 1e6))
ms = timed(lambda: ops.serialize_boards(b, w)); out.append(("othello_serialize_boards", n / ms * 1e3, "positions/s", 80 * n / ms / 1e6))
ms = timed(lambda: ops.learn_accumulate(po)); out.append(("othello_learn_accumulate", positions / ms * 1e3, "positions/s", 16 * positions / ms / 1e6))
for d in (9, 10, 11):
    t0 = time.perf_counter(); nodes = ops.perft(d, device=dev); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); nodes = ops.perft(d, device=dev); dt = min(dt, time.perf_counter() - t0)
    ms = timed(lambda: ops.perft_part(d, 0, 1, device=dev), reps=3)            # the launch sequence alone, result left on the device
    out.append(("othello_perft(%d)=%d" % (d, nodes), nodes / dt, "nodes/s (wall, one synchronous call)", None))
    out.append(("othello_perft_async(%d)" % d, nodes / ms * 1e3, "nodes/s (device time of the launch sequence: %.3f ms)" % ms, None))
small = ops.playout(1 << 16, seed=2, gid0=0, device=dev)
vt = value_table.ValueTable(device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); nrec = vt.update_from_playout(small); torch.cuda.synchronize()
dt = time.perf_counter() - t0
out.append(("value_table.update(65536 games), first batch into an empty table", nrec / dt, "records/s (wall, incl. allocating the table)", None))
dts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); nrec = vt.update_from_playout(small); torch.cuda.synchronize()
    dts.append(time.perf_counter() - t0)
out.append(("value_table.update(65536 games)", nrec / min(dts), "records/s (wall: records + sort + probe + apply)", None))
# the HBM-bound codecs and batch rules at a size where launch latency no longer matters (2^24 positions)
big = 1 << 24
rep = big // n
bb2, ww2 = b.repeat(rep), w.repeat(rep)
t2 = torch.ones(big, dtype=torch.uint8, device=dev)
ms = timed(lambda: ops.serialize_boards(bb2, ww2)); out.append(("othello_serialize_boards @2^24", big / ms * 1e3, "positions/s", 80 * big / ms / 1e6))
chars = ops.serialize_boards(bb2, ww2)
ms = timed(lambda: ops.deserialize_boards(chars)); out.append(("othello_deserialize_boards @2^24", big / ms * 1e3, "positions/s", 80 * big / ms / 1e6))
del chars
ms = timed(lambda: ops.legal(bb2, ww2)); out.append(("othello_legal @2^24", big / ms * 1e3, "positions/s", 24 * big / ms / 1e6))
ms = timed(lambda: ops.features(bb2, ww2, t2)); out.append(("othello_features @2^24", big / ms * 1e3, "positions/s", 57 * big / ms / 1e6))
ms = timed(lambda: ops.evaluate(bb2, ww2, t2, wts)); out.append(("othello_eval @2^24", big / ms * 1e3, "positions/s", 21 * big / ms / 1e6))
ms = timed(lambda: ops.counts(bb2, ww2)); out.append(("othello_counts @2^24", big / ms * 1e3, "positions/s", 28 * big / ms / 1e6))
for name, v, unit, gbs in out:
    print(json.dumps({"kernel": name, "value": v, "unit": unit, "hbm_GBps": gbs}))
