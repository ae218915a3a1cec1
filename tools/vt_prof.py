#!/usr/bin/env python
"""tools/vt_prof.py [log2_games] -- phases of one value-table update (records, stable radix sort, probe, apply)
on a batch of random playouts, timed with CUDA events; records/s of the whole update."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subproc_b200 import _lib, ops, value_table

dev = torch.device('cuda:0')
G = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
po = ops.playout(G, seed=2, gid0=0, device=dev)
L = _lib.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())


def ev(f, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = f(); b.record(); b.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return r, best


out = {"games": G}
vt = value_table.ValueTable(device=dev)
(keys, targets), out["records_ms"] = ev(lambda: vt.records_from_playout(po))
n = keys.numel()
out["records"] = n
_, out["sort_ms"] = ev(lambda: vt.sort_records(keys.clone(), targets.clone()))
_, out["torch_sort_ms"] = ev(lambda: torch.sort(keys, stable=True))
sk, st = vt.sort_records(keys.clone(), targets.clone())
u, c = torch.unique_consecutive(sk, return_counts=True)
out["distinct_keys"], out["longest_run"] = int(u.numel()), int(c.max())
for label in ("first_update", "second_update", "third_update"):      # second: every key already in the table
    k2, t2 = keys.clone(), targets.clone()
    vt.timings = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vt.update(k2, t2)
    torch.cuda.synchronize()
    out[label + "_ms"] = 1e3 * (time.perf_counter() - t0)
    e = vt.timings
    out[label + "_phases_ms"] = dict(zip(("sort", "probe", "reserve", "apply"),
                                         [round(e[i].elapsed_time(e[i + 1]), 3) for i in range(4)]))
    vt.timings = None
# whole update incl. the records kernel, wall clock
vt2 = value_table.ValueTable(device=dev)
vt2.update_from_playout(po)
dts = []
for _ in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vt2.update_from_playout(po)
    torch.cuda.synchronize()
    dts.append(time.perf_counter() - t0)
out["update_from_playout_all_ms"] = [round(1e3 * d, 3) for d in dts]
dt = min(dts)
out["update_from_playout_ms"] = 1e3 * dt
out["records_per_s"] = n / dt
out["table_keys"] = len(vt2)
print(json.dumps(out))
