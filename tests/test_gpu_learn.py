"""Learner statistics kernel vs a numpy restatement over oracle features (B200)."""
import numpy as np
import pytest
import torch

from subproc_b200 import ops
from gpu_util import DEV, host_bits

pytestmark = pytest.mark.gpu


def expected_stats(oracle, po_ref, lam=0.90):
    """x = (mobility, a..h, 1), y = value * lam ** (nplies - t), both sides, per disc-count shard
    (progress_position_moves_learn.py:37-62,112-113) -- in numpy fp64 over oracle features."""
    stats = np.zeros((4, 112))
    n = po_ref['nplies'].size
    for g in range(n):
        L = int(po_ref['nplies'][g])
        b, w = po_ref['black'][:L + 1, g], po_ref['white'][:L + 1, g]
        value = bin(int(po_ref['final_black'][g])).count('1') - bin(int(po_ref['final_white'][g])).count('1')
        for side, sign in ((1, 1.0), (2, -1.0)):
            f = oracle.features(b, w, side).astype(np.float64)
            for t in range(L + 1):
                discs = int(f[t, 0])
                s = 0 if discs <= 16 else 1 if discs <= 32 else 2 if discs <= 48 else 3
                x = np.concatenate([f[t, 1:], [1.0]])
                y = float(sign * value) * (lam ** (L - t))
                stats[s, :100] += np.outer(x, x).reshape(-1)
                stats[s, 100:110] += x * y
                stats[s, 110] += 1
                stats[s, 111] += y * y
    return stats


def test_learn_statistics_match_numpy(oracle):
    n = 300
    po = ops.playout(n, seed=4, gid0=0, device=DEV)
    got = ops.learn_accumulate(po).cpu().numpy()
    want = expected_stats(oracle, oracle.playout(4, 0, n))
    assert np.array_equal(got[:, :100], want[:, :100])                 # integer sums: exact
    assert np.array_equal(got[:, 110], want[:, 110])
    assert np.allclose(got[:, 100:110], want[:, 100:110], rtol=1e-9, atol=1e-9)
    assert np.allclose(got[:, 111], want[:, 111], rtol=1e-9)
    assert got[:, 110].sum() == 2 * (po.total_positions() + n)


def test_learn_statistics_are_additive_over_shards_of_games():
    """the multi-GPU contract: stats(all games) == sum of stats(shards) (what the all-reduce does)"""
    whole = ops.learn_accumulate(ops.playout(4096, seed=8, gid0=0, device=DEV))
    parts = torch.zeros_like(whole)
    for k in range(4):
        ops.learn_accumulate(ops.playout(1024, seed=8, gid0=1024 * k, device=DEV), stats=parts)
    assert torch.equal(whole[:, :100], parts[:, :100])
    assert torch.allclose(whole, parts, rtol=1e-9, atol=1e-9)
