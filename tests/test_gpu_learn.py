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
    acc = ops.learn_accumulate(po)
    got = ops.learn_stats(acc).cpu().numpy()
    want = expected_stats(oracle, oracle.playout(4, 0, n))
    assert np.array_equal(got[:, :100], want[:, :100])                 # integer sums: exact
    assert np.array_equal(got[:, 110], want[:, 110])
    assert np.allclose(got[:, 100:110], want[:, 100:110], rtol=1e-9, atol=1e-9)
    assert np.allclose(got[:, 111], want[:, 111], rtol=1e-9)
    assert got[:, 110].sum() == 2 * (po.total_positions() + n)
    assert np.array_equal(got[:, 109], np.zeros(4))                      # Xty[intercept]: Black's and White's targets cancel
    from subproc_b200 import learner
    assert np.array_equal(learner.stats_from_acc(acc.cpu()), got)        # host and device read the integers alike


def test_learn_statistics_are_additive_over_shards_of_games():
    """the multi-GPU contract: stats(all games) == sum of stats(shards) (what the all-reduce does)"""
    whole = ops.learn_accumulate(ops.playout(4096 + 77, seed=8, gid0=0, device=DEV))
    parts = torch.zeros_like(whole)
    for lo, hi in ((0, 1000), (1000, 1031), (1031, 3000), (3000, 4096 + 77)):        # ragged shards
        ops.learn_accumulate(ops.playout(hi - lo, seed=8, gid0=lo, device=DEV), acc=parts)
    assert torch.equal(whole, parts)                                   # integer accumulators: bit-identical
    assert torch.equal(ops.learn_stats(whole), ops.learn_stats(parts))


def test_device_solve_matches_host_solve():
    """othello_learn_solve (Jacobi on the device) vs learner.solve_shard (numpy eigh on the host) on real
    statistics, incl. an empty shard and a shard whose corner classes are never occupied"""
    from subproc_b200 import learner, parameter
    w0 = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(DEV)
    for n, t_max in ((20000, 120), (3000, 30)):                       # t_max = 30: the last phase shard stays empty
        po = ops.playout(n, seed=14, gid0=0, device=DEV, policy=ops.POLICY_GREEDY, random_plies=8, weights=w0, t_max=t_max)
        if t_max < 120:
            po.nplies.clamp_(max=t_max)                               # treat the truncated prefix as whole games
        stats = ops.learn_stats(ops.learn_accumulate(po))
        w, params, fits = ops.learn_solve(stats, w0)
        host = learner.fit_from_stats(stats)
        f = fits.cpu().numpy()
        for s in range(4):
            if host[s]['n'] == 0:
                assert torch.equal(w[s], w0[s]) and f[s, 12] == 0
                continue
            assert f[s, 12] == host[s]['n']
            assert np.allclose(f[s, :9], host[s]['coef'], rtol=1e-8, atol=1e-10)
            assert abs(f[s, 9] - host[s]['intercept']) <= 1e-8 * max(1.0, abs(host[s]['intercept']))
            assert abs(f[s, 10] - host[s]['rmse']) <= 1e-8 * host[s]['rmse'] and abs(f[s, 11] - host[s]['r2']) < 1e-8
            want = learner.stored_parameters([learner.scale_param(host[s]['coef'])])
            got = params[9 * s:9 * s + 9].cpu().tolist()
            assert max(abs(a - b) for a, b in zip(got, want)) <= 1 and max(abs(v) for v in got) in (126, 127)
            assert w[s, :9].cpu().tolist() == [float(v) for v in got] and float(w[s, 9]) == 0.0


def test_iterations_on_device_track_the_host_loop():
    from subproc_b200 import learner
    A, B = learner.ProgressPositionMovesLearn(), learner.ProgressPositionMovesLearn()
    A.configure({}); B.configure({})
    A.self_play_iteration(8192, seed=5, iteration=0, device=DEV)
    B.self_play_iterations_on_device(1, 8192, seed=5, device=DEV)
    assert max(abs(a - b) for a, b in zip(A.read_parameters(), B.read_parameters())) <= 1
    assert [f['n'] for f in A.last_fits] == [f['n'] for f in B.last_fits]
    B.self_play_iterations_on_device(3, 8192, seed=5, first_iteration=1, device=DEV)
    assert all(-127 <= v <= 127 for v in B.read_parameters()[1:]) and B.last_processed() == 4 * 8192 - 1


def test_refit_in_one_launch_equals_stats_then_solve():
    """othello_learn_refit == othello_learn_stats + othello_learn_solve, bit for bit, and clears the accumulators"""
    from subproc_b200 import parameter
    w0 = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(DEV)
    po = ops.playout(6000, seed=15, gid0=0, device=DEV, policy=ops.POLICY_GREEDY, random_plies=8, weights=w0)
    acc = ops.learn_accumulate(po)
    stats = ops.learn_stats(acc)
    w_a, p_a, f_a = ops.learn_solve(stats, w0)
    stats_b = torch.zeros_like(stats)
    w_b, p_b, f_b = ops.learn_refit(acc.clone(), w0, clear=False, stats_out=stats_b)
    assert torch.equal(stats, stats_b) and torch.equal(w_a, w_b) and torch.equal(p_a, p_b) and torch.equal(f_a, f_b)
    acc2 = acc.clone()
    w_c = w0.clone()
    ops.learn_refit(acc2, w_c, weights_out=w_c, clear=True)              # in place on the weight table
    assert int(acc2.abs().sum()) == 0 and torch.equal(w_c, w_a)
