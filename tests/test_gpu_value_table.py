"""The order-dependent value table on a B200 vs a dict-based restatement of the
reference's loop: __update_state_for_a_book + __update_state_map
(progress_position_moves_learn.py:37-62) driven by oracle features.  Values must be BIT-identical."""
import numpy as np
import pytest
import torch

from subproc_b200 import ops, value_table
from gpu_util import DEV

pytestmark = pytest.mark.gpu


def reference_table_update(oracle, table, ref, a=0.03, l=0.90):
    """books ascending; per book: positions terminal -> start; side 'O' then 'X'"""
    n = ref['nplies'].size
    for g in range(n):
        L = int(ref['nplies'][g])
        b, w = ref['black'][:L + 1, g], ref['white'][:L + 1, g]
        fo, fx = oracle.features(b, w, 1), oracle.features(b, w, 2)
        nb = bin(int(ref['final_black'][g])).count('1')
        nw = bin(int(ref['final_white'][g])).count('1')
        for t in range(L, -1, -1):                                       # book[0] is the terminal record
            for feats, value in ((fo[t], nb - nw), (fx[t], nw - nb)):
                key = tuple(int(v) for v in feats)
                cur = table.get(key, 0.0)                                # `if not exists: set(key, 0)`
                new = float(value) * (l ** (L - t))
                table[key] = new if cur == 0 else cur * (1 - a) + new * a
    return table


def test_value_table_is_bit_identical_to_the_reference_loop(oracle):
    vt = value_table.ValueTable(device=DEV)
    want = {}
    for batch, (seed, gid0, n) in enumerate(((31, 0, 150), (31, 150, 90))):       # two batches: table carries over
        po = ops.playout(n, seed=seed, gid0=gid0, device=DEV)
        nrec = vt.update_from_playout(po)
        ref = oracle.playout(seed, gid0, n)
        assert nrec == 2 * int((ref['nplies'] + 1).sum())
        reference_table_update(oracle, want, ref)
        got = vt.items()
        assert set(got) == set(want)
        bad = [k for k in want if got[k] != want[k]]
        assert not bad, (batch, bad[:3], [(got[k], want[k]) for k in bad[:3]])
    # the start position is visited by every game, twice: the longest sequential run
    assert vt.get((4, 4, 0, 0, 0, 0, 0, 0, 0, 2)) == want[(4, 4, 0, 0, 0, 0, 0, 0, 0, 2)]
    assert value_table.unpack_key(value_table.pack_key((64, 33, 4, 8, 4, 8, 8, 16, 4, 12))) == (64, 33, 4, 8, 4, 8, 8, 16, 4, 12)


def test_table_fit_runs_and_scales_like_the_reference():
    vt = value_table.ValueTable(device=DEV)
    vt.update_from_playout(ops.playout(20000, seed=32, gid0=0, device=DEV))
    assert len(vt) > 100000
    mse, score, param, nsample = vt.fit_parameter(33, 48, num=20000, seed=1)
    assert nsample >= 20000 and len(param) == 9 and abs(max(abs(p) for p in param) - 127) < 1e-9
    assert mse > 0 and score <= 1
    f = vt.features().cpu().numpy()
    assert f[:, 0].min() >= 4 and f[:, 0].max() <= 64 and (f[:, 2:].sum(axis=1) <= f[:, 0]).all()


def test_book_driven_learn_and_update_batch_equals_trajectory_path(oracle):
    """the reference's entry point (books in the recorder schema, reversed, terminal first) must build
    the same table as the trajectory path, and refit parameters from it"""
    from subproc_b200 import books, learner
    n = 40
    po = ops.playout(n, seed=33, gid0=0, device=DEV)
    vt = value_table.ValueTable(device=DEV)
    vt.update_from_playout(po)
    bks = []
    for i, (recs, meta) in enumerate(books.books_from_playout(po)):
        rev = list(reversed(sorted(recs, key=lambda r: int(r['turn']))))        # learn_books, replearn.py:37-38
        assert rev[0]['end']
        bks.append((i + 1, rev, meta))
    L = learner.ProgressPositionMovesLearn()
    L.configure({})
    mses, scores, params, nsamples = L.learn_and_update_batch(bks, device=DEV, sample=2000)
    assert L.table.items() == vt.items()                                        # bit-identical values, same keys
    assert L.last_processed() == n and len(params) == 4 and all(len(p) == 9 for p in params)
    rp = L.read_parameters()
    assert rp[0] == 2 and len(rp) == 37 and all(-127 <= v <= 127 for v in rp[1:])
    stats = L.store_batch_stats(po)
    c = po.final_counts().cpu().numpy()
    assert stats['min_disc_diff'] == int((c[:, 0] - c[:, 1]).min()) and len(stats['diffs']) == n


def test_learner_checkpoint_round_trip(tmp_path):
    """resume = the reference restarting on its Redis state: last_processed, parameters, value table"""
    from subproc_b200 import books, learner
    po = ops.playout(30, seed=35, gid0=0, device=DEV)
    bks = [(i + 1, list(reversed(recs)), meta) for i, (recs, meta) in enumerate(books.books_from_playout(po))]
    A = learner.ProgressPositionMovesLearn(); A.configure({})
    A.learn_and_update_batch(bks[:20], device=DEV, sample=500)
    A.save(str(tmp_path / "learner.pt"))
    B = learner.ProgressPositionMovesLearn().load(str(tmp_path / "learner.pt"), device=DEV)
    assert B.last_processed() == 20 and B.read_parameters() == A.read_parameters() and B.table.items() == A.table.items()
    A.learn_and_update_batch(bks[20:], device=DEV, sample=500, seed=9)
    B.learn_and_update_batch(bks[20:], device=DEV, sample=500, seed=9)
    assert B.table.items() == A.table.items() and B.read_parameters() == A.read_parameters()
