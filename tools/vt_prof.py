import sys, time, torch
sys.path.insert(0, '/root/repo')
from subproc_b200 import ops, value_table
dev = torch.device('cuda:0')
po = ops.playout(1 << 16, seed=2, gid0=0, device=dev)
vt = value_table.ValueTable(device=dev)
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
for rep in range(2):
    (keys, targets), ms = T(lambda: vt.records_from_playout(po)); print('records', ms, keys.numel())
    (sk, perm), ms = T(lambda: torch.sort(keys, stable=True)); print('sort', ms)
    (u, c), ms = T(lambda: torch.unique_consecutive(sk, return_counts=True)); print('unique', ms, u.numel(), int(c.max()))
    vt2 = value_table.ValueTable(device=dev)
    _, ms = T(lambda: vt2.update(keys, targets)); print('update total', ms)
    _, ms = T(lambda: vt2.update(keys, targets)); print('update again (merge)', ms)
