"""Host side of the learner: regression from sufficient statistics vs scikit-learn (the library the
reference calls, progress_position_moves_learn.py:167-181), parameter scaling / truncation, the
paramgen wire format (golden bytes written by the reference's own paramgen.write_data), and the
world_size-2 all-reduce path over gloo."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden
from subproc_b200 import learner, paramgen, parameter


def stats_from_samples(x9, y, discs):
    """what the learn kernel accumulates, in numpy: x = (9 features, 1)"""
    stats = np.zeros((4, 112))
    x = np.concatenate([x9, np.ones((x9.shape[0], 1))], axis=1)
    for s, (lo, hi) in enumerate(learner.PHASE_SHARDS):
        m = (discs >= lo) & (discs <= hi)
        xs, ys = x[m], y[m]
        stats[s, :100] = (xs.T @ xs).reshape(-1)
        stats[s, 100:110] = xs.T @ ys
        stats[s, 110] = m.sum()
        stats[s, 111] = (ys * ys).sum()
    return stats


def synthetic(n, seed, dead_columns=()):
    rng = np.random.RandomState(seed)
    x9 = rng.randint(0, 12, size=(n, 9)).astype(np.float64)
    for c in dead_columns:
        x9[:, c] = 0
    w = rng.uniform(-3, 3, size=9)
    y = x9 @ w + 1.5 + rng.normal(0, 2.0, size=n)
    discs = rng.randint(4, 65, size=n)
    return x9, y, discs


def test_solve_matches_sklearn_within_1e8():
    from sklearn import linear_model
    x9, y, discs = synthetic(20000, 1)
    fits = learner.fit_from_stats(stats_from_samples(x9, y, discs))
    for s, (lo, hi) in enumerate(learner.PHASE_SHARDS):
        m = (discs >= lo) & (discs <= hi)
        lr = linear_model.LinearRegression(fit_intercept=True).fit(x9[m], y[m])
        assert np.allclose(fits[s]['coef'], lr.coef_, rtol=1e-8, atol=1e-10)          # 1e-8 rel
        assert abs(fits[s]['intercept'] - lr.intercept_) <= 1e-8 * max(1, abs(lr.intercept_))
        assert abs(fits[s]['rmse'] - np.sqrt(np.mean((lr.predict(x9[m]) - y[m]) ** 2))) < 1e-8
        assert abs(fits[s]['r2'] - lr.score(x9[m], y[m])) < 1e-8
        assert fits[s]['n'] == m.sum()


def test_rank_deficient_shard_gets_minimum_norm_solution_like_sklearn():
    from sklearn import linear_model
    x9, y, discs = synthetic(5000, 2, dead_columns=(1, 2, 3))     # corner classes empty early in the game
    discs[:] = 10
    fit = learner.fit_from_stats(stats_from_samples(x9, y, discs))[0]
    lr = linear_model.LinearRegression(fit_intercept=True).fit(x9, y)
    assert np.allclose(fit['coef'], lr.coef_, rtol=1e-8, atol=1e-9)
    assert np.all(fit['coef'][[1, 2, 3]] == 0)
    empty = learner.fit_from_stats(np.zeros((4, 112)))[2]
    assert empty['n'] == 0 and np.all(empty['coef'] == 0)


def test_scaling_and_truncation():
    coef = np.array([0.5, -2.0, 1.0, 0, 0, 0, 0, 0, 0.25])
    p = learner.scale_param(coef)
    assert p == tuple(float(c) * (127 / 2.0) for c in coef)        # progress_position_moves_learn.py:180-181
    assert learner.stored_parameters([p]) == [31, -127, 63, 0, 0, 0, 0, 0, 15]   # int() truncates toward zero (:200)
    L = learner.ProgressPositionMovesLearn()
    L.configure({})
    assert L.name() == 'progresspositionmovelearn' and (L.a, L.b, L.l) == (0.03, 0.003, 0.90)
    rp = L.read_parameters()
    assert len(rp) == 37 and rp[0] == 2 and list(rp[1:10]) == [100, 99, -1, -1, -1, -1, 3, 8, 20]
    assert L.weights_table().shape == (4, 10) and L.weights_table()[3, 1] == 100


def test_paramgen_wire_format_golden():
    for case in load_golden("paramgen.json.gz"):
        data = paramgen.encode(case['params'])
        assert data.hex() == case['bytes'] and len(data) == 38 and data[-1] == 0
        assert list(paramgen.decode(data)) == case['params']


def test_parameter_hash_format_and_book_strings():
    P = parameter.ProgressPositionMovesParameter()
    key = 'othelloparam:progresspositionmovelearn:param:state:12:5:1:0:2:3:0:4:1:2'
    assert P.phase_from_hash(key) == 12 and P.features_from_hash(key) == (5, 1, 0, 2, 3, 0, 4, 1, 2)
    b, w = parameter.bits_from_book_string('---------------------------XO------OX---------------------------')
    assert (b, w) == (0x0000000810000000, 0x0000001008000000)
    assert parameter.ProgressPositionMovesParameter().header() == 2


def test_shard_of_games():
    assert [learner.shard_of_games(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert learner.shard_of_games(1 << 22, 7, 8) == (7 << 19, 1 << 22)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x9, y, discs = synthetic(6000, 3)
    lo, hi = learner.shard_of_games(6000, rank, world)
    stats = torch.from_numpy(stats_from_samples(x9[lo:hi], y[lo:hi], discs[lo:hi]))
    L = learner.ProgressPositionMovesLearn()
    rows = L.learn_from_stats(stats)
    q.put((rank, L.read_parameters(), [list(r) for r in rows]))
    dist.destroy_process_group()


def test_world_size_2_allreduce_equals_single_process_fit():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x9, y, discs = synthetic(6000, 3)
    L = learner.ProgressPositionMovesLearn()
    import torch
    rows = L.learn_from_stats(torch.from_numpy(stats_from_samples(x9, y, discs)))
    for rank, params, rrows in got:
        assert np.allclose(np.array(rrows), np.array(rows), rtol=1e-9)     # tolerance 1e-9
        # int() truncation (:200) may differ by one only where the float sits on an integer boundary
        diff = np.abs(np.array(params) - np.array(L.read_parameters()))
        assert diff.max() <= 1
        flat = np.array(rows).reshape(-1)
        assert all(abs(flat[i] - round(flat[i])) < 1e-6 for i in np.flatnonzero(diff[1:]))
    assert got[0][1] == got[1][1]                                          # every rank ends with identical parameters


def test_batch_stats_payload_follows_the_reference_loop():
    rng = np.random.RandomState(8)
    b = rng.randint(0, 65, size=300)
    w = 64 - b - rng.randint(0, 3, size=300)
    w = np.maximum(w, 0)
    # learn_base.py:58-98, restated literally
    black_wins = white_wins = 0
    diff = []
    for bd, wd in zip(b, w):
        diff.append(int(bd - wd))
        if bd > wd:
            black_wins += 1
        elif wd > black_wins:
            white_wins += 1
    got = learner.batch_stats(b, w, 'A', 'B', 'p')
    assert got['A_win_rate'] == float(black_wins) / 300 and got['B_win_rate'] == float(white_wins) / 300
    assert got['min_disc_diff'] == min(diff) and got['max_disc_diff'] == max(diff)
    assert got['avg_disc_diff'] == float(sum(diff)) / 300 and got['diffs'] == sorted(diff) and got['params_used'] == 'p'
    assert got['white_wins_by_discs'] == int((w > b).sum())


def acc_from_samples(x9, y, discs):
    """the exact integer accumulators of othello_learn_accumulate (include/othello_b200.h) in numpy /
    Python integers, every sample handed over as its own 2^-40 fixed-point partial sum"""
    acc = np.zeros((4, learner.N_ACC), dtype=np.int64)
    x = np.concatenate([x9, np.ones((x9.shape[0], 1))], axis=1).astype(np.int64)
    for s, (lo, hi) in enumerate(learner.PHASE_SHARDS):
        m = (discs >= lo) & (discs <= hi)
        g = x[m].T @ x[m]
        for i in range(10):
            for j in range(i, 10):
                acc[s, learner._pair(i, j)] = g[i, j]
        for k in range(10):
            vals = (x[m][:, k] * y[m]) if k < 9 else (y[m] * y[m])
            q = sum(int(round(float(v) * 2.0 ** 40)) for v in vals)
            acc[s, 56 + 2 * k], acc[s, 57 + 2 * k] = q >> 32, q & 0xFFFFFFFF
    return acc


def test_stats_from_acc_reads_the_integers_exactly():
    x9, y, discs = synthetic(3000, 5)
    y = np.round(y * 64) / 64                                     # dyadic targets: the fixed point holds them exactly
    got = learner.stats_from_acc(acc_from_samples(x9, y, discs))
    want = stats_from_samples(x9, y, discs)
    assert np.array_equal(got[:, :100], want[:, :100]) and np.array_equal(got[:, 110], want[:, 110])
    assert np.allclose(got[:, 100:109], want[:, 100:109], rtol=1e-13) and np.allclose(got[:, 111], want[:, 111], rtol=1e-13)
    assert np.array_equal(got[:, 109], np.zeros(4))               # the kernel never accumulates Xty[intercept]


def _acc_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x9, y, discs = synthetic(6000, 4)
    lo, hi = learner.shard_of_games(6000, rank, world)
    L = learner.ProgressPositionMovesLearn()
    L.learn_from_acc(torch.from_numpy(acc_from_samples(x9[lo:hi], y[lo:hi], discs[lo:hi])))
    q.put((rank, L.read_parameters(), [f['coef'].tolist() for f in L.last_fits]))
    dist.destroy_process_group()


def test_world_size_2_allreduce_of_integer_accumulators_is_exact():
    """the N > 1 learner path over gloo: all-reduce of the int64 accumulators, then the same bits as one process"""
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_acc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x9, y, discs = synthetic(6000, 4)
    L = learner.ProgressPositionMovesLearn()
    L.learn_from_acc(torch.from_numpy(acc_from_samples(x9, y, discs)))
    for rank, params, coefs in got:
        assert params == L.read_parameters()                       # == , not a tolerance
        assert coefs == [f['coef'].tolist() for f in L.last_fits]
