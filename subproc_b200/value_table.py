"""The reference's order-dependent value table on the GPU, with its exact update semantics.

``ValueTable.update_from_playout`` does what ``learn_and_update_batch`` does book by book through
Redis (progress_position_moves_learn.py:37-62,88-91): every (position, side) of every game updates the
float stored under its ``counts()`` 10-tuple, in the reference's order.  The result is bit-identical to
running the reference's loop (tests/test_gpu_value_table.py checks it against vectors produced by
executing the reference's own update methods, tests/golden/value_table.json.gz); the table lives in HBM
as a hash from key to a dense (key, value) array instead of Redis strings.

``fit_parameter`` is the reference's per-shard fit on a sample of the TABLE (:66-86,160-184): draw
keys at random, keep those whose disc count is in the shard, de-duplicate, bootstrap-resample, OLS with
intercept, RMSE / R^2 on a second sample, scale to +-127.  The reference draws with Python's unseeded
``random`` and sklearn's ``resample``; here a seeded ``torch.Generator`` stands in (the distribution is
the same, the draws are not -- nothing in the reference pins them).
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib, ops, learner

WIDTHS = (7, 6, 3, 4, 3, 4, 4, 5, 3, 4)          # discs mobility a b c d e f g h -> 43 bits
KEY_BITS = sum(WIDTHS)


def pack_key(features):
    k = 0
    for v, w in zip(features, WIDTHS):
        k = (k << w) | int(v)
    return k


def unpack_key(k):
    out = []
    for w in reversed(WIDTHS):
        out.append(k & ((1 << w) - 1))
        k >>= w
    return tuple(reversed(out))


class ValueTable(object):
    """The table in HBM: dense ``keys`` / ``values`` arrays (in order of first appearance: batch by batch,
    ascending key inside a batch) plus an open-addressing hash key -> dense index.  Everything on the update
    path is a hand-written kernel of csrc/value_table.cu: records, stable radix sort, probe, apply."""

    MIN_LOG2_CAPACITY = 12

    def __init__(self, device=None, a=0.03, lam=0.90):
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.a = a
        self.lam = lam
        self.n = 0                                             # keys in the table
        self._keys = torch.empty(0, dtype=torch.int64, device=self.device)
        self._values = torch.empty(0, dtype=torch.float64, device=self.device)
        self._log2cap = self.MIN_LOG2_CAPACITY
        self._slot_keys = torch.zeros(1 << self._log2cap, dtype=torch.int64, device=self.device)
        self._slot_idx = torch.zeros(1 << self._log2cap, dtype=torch.int32, device=self.device)
        self._counters = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._ws = {}
        self.timings = None                                    # set to a list to collect CUDA events of update()

    def __len__(self):
        return self.n

    @property
    def keys(self):
        return self._keys[:self.n]

    @property
    def values(self):
        return self._values[:self.n]

    def load_state(self, keys, values):
        """adopt dense key / value arrays (a checkpoint) and rebuild the hash"""
        self.n = int(keys.numel())
        self._keys = keys.to(self.device, torch.int64).contiguous().clone()
        self._values = values.to(self.device, torch.float64).contiguous().clone()
        self._log2cap = self.MIN_LOG2_CAPACITY
        self._reserve(self.n, force=True)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _scratch(self, name, nbytes):
        """grow-only device scratch, reused by every update"""
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes:
            t = self._ws[name] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return t

    def _reserve(self, n_total, force=False):
        """room for n_total keys: dense arrays grow geometrically, the hash keeps a load factor <= 1/2"""
        if self._keys.numel() < n_total:
            cap = max(n_total, 2 * self._keys.numel(), 1024)
            for name in ("_keys", "_values"):
                old = getattr(self, name)
                new = torch.empty(cap, dtype=old.dtype, device=self.device)
                new[:self.n] = old[:self.n]
                setattr(self, name, new)
        log2 = self._log2cap
        while (1 << log2) < 2 * max(n_total, 1):
            log2 += 1
        if log2 != self._log2cap or force:
            self._log2cap = log2
            self._slot_keys = torch.zeros(1 << log2, dtype=torch.int64, device=self.device)
            self._slot_idx = torch.zeros(1 << log2, dtype=torch.int32, device=self.device)
            P = lambda t: ctypes.c_void_p(t.data_ptr())
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().othello_table_rehash(P(self._keys), self.n, P(self._slot_keys), P(self._slot_idx),
                                                           log2, self._stream()), "othello_table_rehash")

    # ---- update ------------------------------------------------------------------------------
    def records_from_playout(self, po):
        """(keys, targets) of every (position, side) in the reference's update order."""
        n = po.n_games
        counts = 2 * (po.nplies.to(torch.int64) + 1)
        counts = torch.where(po.nplies > po.t_max, torch.zeros_like(counts), counts)     # truncated games are skipped
        base = torch.cumsum(counts, 0) - counts
        total = int(counts.sum().item())
        keys = torch.empty(total, dtype=torch.int64, device=self.device)
        targets = torch.empty(total, dtype=torch.float64, device=self.device)
        decay = ops.decay_tensor(po.t_max, self.lam, self.device)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().othello_value_records(
                P(po.black), P(po.white), P(po.nplies), P(po.final_black), P(po.final_white), n, n, po.t_max,
                P(decay), P(base), P(keys), P(targets), self._stream()), "othello_value_records")
        return keys, targets

    def sort_records(self, keys, targets):
        """stable sort by key IN PLACE (othello_sort_records): groups a key's records, keeps their order"""
        n = keys.numel()
        if n > 1:
            L = _lib.lib()
            kalt, valt = torch.empty_like(keys), torch.empty_like(targets)
            nbytes = int(L.othello_sort_workspace_bytes(n))
            ws = self._scratch("sort", nbytes)
            P = lambda t: ctypes.c_void_p(t.data_ptr())
            with torch.cuda.device(self.device):
                _lib.check(L.othello_sort_records(P(keys), P(targets), P(kalt), P(valt), n, KEY_BITS, P(ws), nbytes,
                                                  self._stream()), "othello_sort_records")
        return keys, targets

    def update(self, keys, targets):
        """apply records (in update order) to the table; ``keys`` / ``targets`` are sorted in place"""
        n = keys.numel()
        if n == 0:
            return
        L = _lib.lib()
        tm = self.timings
        ev = (lambda: None) if tm is None else (lambda: tm.append(self._event()))
        ev()
        self.sort_records(keys, targets)
        ev()
        nbytes = int(L.othello_table_workspace_bytes(n))
        ws = self._scratch("table", nbytes)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(L.othello_table_probe(P(keys), n, P(self._slot_keys), P(self._slot_idx), self._log2cap, P(ws), nbytes,
                                             P(self._counters), self._stream()), "othello_table_probe")
            ev()
            n_new = int(self._counters[1].item())                 # the one host round trip of an update: sizing
            self._reserve(self.n + n_new)
            ev()
            _lib.check(L.othello_table_apply(P(keys), P(targets), n, self.a, P(self._slot_keys), P(self._slot_idx),
                                             self._log2cap, P(self._keys), P(self._values), self.n, P(ws), self._stream()),
                       "othello_table_apply")
            ev()
        self.n += n_new

    def _event(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        return e

    def update_from_playout(self, po):
        keys, targets = self.records_from_playout(po)
        self.update(keys, targets)
        return keys.numel()

    def update_sharded(self, keys, targets):
        """One table over all ranks of the default process group: rank r owns the keys with
        owner_of(key) == r.  Every rank hands in ITS records in its own update order; ranks must hold
        contiguous, ascending blocks of game ids (rank 0 the lowest), so that records received in rank order
        are in the global update order.  One all_to_all of (key, target) records per batch; afterwards this
        rank's table holds its share of the keys with exactly the values of a single-GPU table."""
        import torch.distributed as dist
        world = dist.get_world_size()
        L = _lib.lib()
        n = keys.numel()
        nbytes = int(L.othello_sort_workspace_bytes(n))
        ws = self._scratch("sort", nbytes)
        kout, vout = torch.empty_like(keys), torch.empty_like(targets)
        counts = torch.zeros(world, dtype=torch.int64, device=self.device)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(L.othello_partition_records(P(keys), P(targets), P(kout), P(vout), n, world, P(counts), P(ws), nbytes,
                                                   self._stream()), "othello_partition_records")
        incoming = torch.empty_like(counts)
        dist.all_to_all_single(incoming, counts)
        send, recv = counts.cpu().tolist(), incoming.cpu().tolist()
        rk = torch.empty(sum(recv), dtype=torch.int64, device=self.device)
        rv = torch.empty(sum(recv), dtype=torch.float64, device=self.device)
        dist.all_to_all_single(rk, kout, recv, send)
        dist.all_to_all_single(rv, vout, recv, send)
        self.update(rk, rv)                                       # blocks arrive in rank order = update order
        return rk.numel()

    # ---- read --------------------------------------------------------------------------------
    def lookup(self, keys):
        """dense index of every key (int64 tensor) or -1"""
        out = torch.empty(keys.numel(), dtype=torch.int32, device=self.device)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().othello_table_lookup(P(keys), keys.numel(), P(self._slot_keys), P(self._slot_idx),
                                                       self._log2cap, P(out), self._stream()), "othello_table_lookup")
        return out

    def get(self, features):
        """value stored under a counts() 10-tuple, or None"""
        if not len(self):
            return None
        at = int(self.lookup(torch.tensor([pack_key(features)], dtype=torch.int64, device=self.device)).item())
        return float(self._values[at].item()) if at >= 0 else None

    def features(self, keys=None):
        keys = (self.keys if keys is None else keys).contiguous()
        out = torch.empty((keys.numel(), 10), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().othello_unpack_keys(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                      keys.numel(), self._stream()), "othello_unpack_keys")
        return out

    def items(self):
        f = self.features().cpu().numpy()
        v = self.values.cpu().numpy()
        return {tuple(int(x) for x in f[i]): float(v[i]) for i in range(len(v))}

    # ---- the reference's fit on a sample of the table ------------------------------------------
    def random_sample(self, num, p_min, p_max, generator):
        """__random_sample (:66-86): up to 5 rounds of `num` uniform draws over ALL keys, keep those in
        the shard, de-duplicate; returns (x int32 [m][9], y float64 [m])."""
        n = len(self)
        if n == 0:
            return (torch.empty((0, 9), dtype=torch.int32, device=self.device),
                    torch.empty(0, dtype=torch.float64, device=self.device))
        discs = self.keys >> 36                                        # top 7 bits of the 43-bit key
        chosen = torch.zeros(n, dtype=torch.bool, device=self.device)
        for _ in range(5):
            idx = torch.randint(0, n, (num,), generator=generator, device=self.device)
            ok = (discs[idx] >= p_min) & (discs[idx] <= p_max)
            chosen[idx[ok]] = True
            if int(chosen.sum().item()) >= num:
                break
        sel = torch.nonzero(chosen).reshape(-1)
        return self.features(self.keys[sel].contiguous())[:, 1:].contiguous(), self.values[sel]

    def fit_parameter(self, phase_from, phase_to, num=50000, seed=0):
        """(mse, score, param, nsample) like the reference's worker job (:160-184)"""
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed)
        x, y = self.random_sample(num, phase_from, phase_to, gen)
        m = y.numel()
        if m == 0:
            return float('nan'), float('nan'), tuple([0.0] * 9), 0
        boot = torch.randint(0, m, (m,), generator=gen, device=self.device)       # sklearn.utils.resample
        fit = learner.solve_shard(_stats_row(x[boot], y[boot]))
        tx, ty = self.random_sample(num, phase_from, phase_to, gen)
        pred = tx.to(torch.float64) @ torch.from_numpy(fit['coef']).to(self.device) + fit['intercept']
        mse = math.sqrt(float(((pred - ty) ** 2).mean().item()))
        sst = float(((ty - ty.mean()) ** 2).sum().item())
        score = 1.0 - float(((pred - ty) ** 2).sum().item()) / sst if sst > 0 else float('nan')
        return mse, score, learner.scale_param(fit['coef']), m


def _stats_row(x, y):
    """normal-equation statistics [112] of a sample (x int32 [m][9], y float64 [m])"""
    xd = torch.cat([x.to(torch.float64), torch.ones((x.shape[0], 1), dtype=torch.float64, device=x.device)], dim=1)
    row = torch.zeros(112, dtype=torch.float64, device=x.device)
    row[:100] = (xd.t() @ xd).reshape(-1)
    row[100:110] = xd.t() @ y
    row[110] = x.shape[0]
    row[111] = (y * y).sum()
    return row.cpu().numpy()
