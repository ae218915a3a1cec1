"""The reference's order-dependent value table on the GPU, with its exact update semantics.

``ValueTable.update_from_playout`` does what ``learn_and_update_batch`` does book by book through
Redis (progress_position_moves_learn.py:37-62,88-91): every (position, side) of every game updates the
float stored under its ``counts()`` 10-tuple, in the reference's order.  The result is bit-identical to
running the reference's loop (tests/test_gpu_value_table.py checks it against a dict-based
restatement); the table lives in HBM as a sorted key array + value array instead of Redis strings.

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
    def __init__(self, device=None, a=0.03, lam=0.90):
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.a = a
        self.lam = lam
        self.keys = torch.empty(0, dtype=torch.int64, device=self.device)      # sorted, unique
        self.values = torch.empty(0, dtype=torch.float64, device=self.device)

    def __len__(self):
        return int(self.keys.numel())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

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
        decay = torch.from_numpy(ops.decay_table(po.t_max, self.lam)).to(self.device)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().othello_value_records(
                P(po.black), P(po.white), P(po.nplies), P(po.final_black), P(po.final_white), n, n, po.t_max,
                P(decay), P(base), P(keys), P(targets), self._stream()), "othello_value_records")
        return keys, targets

    def update(self, keys, targets):
        """apply records (already in update order) to the table"""
        if keys.numel() == 0:
            return
        skeys, perm = torch.sort(keys, stable=True)                    # groups keys, keeps the order inside a key
        stargets = targets[perm]
        uniq, cnt = torch.unique_consecutive(skeys, return_counts=True)
        seg = torch.zeros(uniq.numel() + 1, dtype=torch.int64, device=self.device)
        seg[1:] = torch.cumsum(cnt, 0)
        init = torch.zeros(uniq.numel(), dtype=torch.float64, device=self.device)
        if self.keys.numel():
            pos = torch.searchsorted(self.keys, uniq).clamp_(max=self.keys.numel() - 1)
            hit = self.keys[pos] == uniq
            init[hit] = self.values[pos[hit]]
        out = torch.empty_like(init)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().othello_value_smooth(P(stargets), P(seg), P(init), self.a, P(out), uniq.numel(),
                                                       self._stream()), "othello_value_smooth")
        if self.keys.numel():
            keep = torch.ones(self.keys.numel(), dtype=torch.bool, device=self.device)
            keep[pos[hit]] = False
            allk = torch.cat([self.keys[keep], uniq])
            allv = torch.cat([self.values[keep], out])
            order = torch.argsort(allk)
            self.keys, self.values = allk[order], allv[order]
        else:
            self.keys, self.values = uniq, out

    def update_from_playout(self, po):
        keys, targets = self.records_from_playout(po)
        self.update(keys, targets)
        return keys.numel()

    # ---- read --------------------------------------------------------------------------------
    def get(self, features):
        """value stored under a counts() 10-tuple, or None"""
        k = pack_key(features)
        if not len(self):
            return None
        pos = int(torch.searchsorted(self.keys, torch.tensor([k], dtype=torch.int64, device=self.device)).item())
        if pos < len(self) and int(self.keys[pos].item()) == k:
            return float(self.values[pos].item())
        return None

    def features(self, keys=None):
        keys = self.keys if keys is None else keys
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
