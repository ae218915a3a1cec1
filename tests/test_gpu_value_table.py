"""The order-dependent value table on a B200.  Pinned twice: to vectors produced by
EXECUTING the reference's own __update_state_for_a_book / __update_state_map
(progress_position_moves_learn.py:37-62; tests/golden/value_table.json.gz, oracle/make_golden.py) and, on
other games, to a dict-based restatement of that loop driven by oracle features.  Values must be
BIT-identical."""
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


def test_value_table_equals_what_the_reference_text_computes():
    """golden: the reference's own update methods, run from their source text on the first 24 golden games
    in two batches (the table carries over); the CUDA table must hold the same keys and the same bits"""
    from conftest import load_golden
    gold = load_golden("value_table.json.gz")
    vt = value_table.ValueTable(device=DEV, a=gold['a'], lam=gold['l'])
    for (lo, hi), want in zip(gold['batches'], gold['after']):
        vt.update_from_playout(ops.playout(hi - lo, seed=gold['seed'], gid0=lo, device=DEV))
        got = {':'.join(str(v) for v in k): val.hex() for k, val in vt.items().items()}
        assert got == want


def test_stable_radix_sort_and_owner_partition():
    """othello_sort_records == a stable sort by key (order inside a key preserved), for ragged sizes and
    heavy duplicates; othello_partition_records == a stable split by owner whose blocks add up"""
    import ctypes
    from subproc_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device=DEV); g.manual_seed(5)
    vt = value_table.ValueTable(device=DEV)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for n, distinct in ((1, 1), (4095, 7), (4097, 300), (200000 + 13, 1 << 40), (1 << 20, 1000)):
        keys = torch.randint(1, distinct + 1, (n,), generator=g, device=DEV, dtype=torch.int64)
        keys |= torch.randint(4, 65, (n,), generator=g, device=DEV, dtype=torch.int64) << 36       # disc count on top
        vals = torch.arange(n, device=DEV, dtype=torch.float64)                                    # = original position
        want_k, perm = torch.sort(keys, stable=True)
        k2, v2 = vt.sort_records(keys.clone(), vals.clone())
        assert torch.equal(k2, want_k) and torch.equal(v2, vals[perm])
        for world in (2, 8):
            ko, vo = torch.empty_like(keys), torch.empty_like(vals)
            counts = torch.zeros(world, dtype=torch.int64, device=DEV)
            nbytes = int(L.othello_sort_workspace_bytes(n))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
            assert L.othello_partition_records(P(keys), P(vals), P(ko), P(vo), n, world, P(counts), P(ws), nbytes, None) == 0
            c = counts.cpu().tolist()
            assert sum(c) == n
            at = 0
            owners = {}
            for r in range(world):
                blk_k, blk_v = ko[at:at + c[r]], vo[at:at + c[r]]
                assert bool((blk_v[1:] > blk_v[:-1]).all())                # order preserved inside an owner
                assert torch.equal(keys[blk_v.long()], blk_k)
                for k in blk_k[:50].cpu().tolist():
                    assert owners.setdefault(k, r) == r                   # a key has one owner
                at += c[r]


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


def test_self_play_iteration_with_the_references_table_semantics():
    """greedy self-play -> exact value table -> per-shard fit on samples of the table, iterated: the table carries
    over, the parameters stay in the stored range, and the table equals one built from the same games directly"""
    from subproc_b200 import learner
    L = learner.ProgressPositionMovesLearn(); L.configure({})
    po0, (mses, scores, params, nsamples) = L.self_play_iteration_table(3000, seed=41, iteration=0, device=DEV, sample=5000)
    assert len(params) == 4 and all(len(p) == 9 for p in params) and sum(nsamples) > 0
    rp0 = L.read_parameters()
    assert rp0[0] == 2 and all(-127 <= v <= 127 for v in rp0[1:])
    n0 = len(L.table)
    po1, _ = L.self_play_iteration_table(3000, seed=41, iteration=1, device=DEV, sample=5000)
    assert len(L.table) > n0 and L.last_processed() == 5999
    vt = value_table.ValueTable(device=DEV)
    vt.update_from_playout(po0); vt.update_from_playout(po1)
    assert vt.items() == L.table.items()
