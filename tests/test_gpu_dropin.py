"""Drop-in behaviour on a B200: the reference's own callers run against subproc_b200's modules.

Where the reference sources travel with the box (oracle/_ref, the py3 transcription written by
oracle/build_ref.py) the reference's parameter_progress_position_moves_learn.py is executed
UNMODIFIED on top of ``subproc_b200.board`` registered as module ``board`` -- the drop-in the
north_star asks for.  Everything is also checked against the golden vectors."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from subproc_b200 import board, parameter, learner, ops
from gpu_util import DEV, h

pytestmark = pytest.mark.gpu


def test_counts_and_hash_match_golden_features(golden_games):
    P = parameter.ProgressPositionMovesParameter()
    g = golden_games[0]
    books = [{'book': p['ser'][:64], 'whosturn': p['ser'][65], 'turn': p['nturn']} for p in g['positions']]
    for bk, p in list(zip(books, g['positions']))[::7]:
        assert list(parameter.counts(bk, 'O')) == p['feat_O']
        assert list(parameter.counts(bk, 'X')) == p['feat_X']
        assert P.hash_from_book(bk, 'O') == ':'.join(str(v) for v in p['feat_O'])
    hs = P.hashes_from_books(books, ['X'] * len(books))
    assert hs == [':'.join(str(v) for v in p['feat_X']) for p in g['positions']]
    bb = parameter.board_from_a_book(books[3])
    assert ("%016x" % bb._black, "%016x" % bb._white, bb.turn, bb.nturn) == \
        (g['positions'][3]['b'], g['positions'][3]['w'], g['positions'][3]['turn'], g['positions'][3]['nturn'])


def test_reference_parameter_module_runs_on_our_board(golden_games):
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "parameter.py")):
        pytest.skip("reference sources not on this box")
    saved = {k: sys.modules.get(k) for k in ("board", "parameter", "parameter_progress_position_moves_learn")}
    try:
        sys.modules["board"] = board                                  # the drop-in
        mods = {}
        for name in ("parameter", "parameter_progress_position_moves_learn"):
            m = types.ModuleType(name)
            sys.modules[name] = m
            exec(compile(open(os.path.join(ref_dir, name + ".py")).read(), name, "exec"), m.__dict__)
            mods[name] = m
        ref_counts = mods["parameter_progress_position_moves_learn"].counts
        g = golden_games[1]
        for p in g['positions'][::9]:
            bk = {'book': p['ser'][:64], 'whosturn': p['ser'][65], 'turn': p['nturn']}
            assert list(ref_counts(bk, 'O')) == list(parameter.counts(bk, 'O'))
            assert list(ref_counts(bk, 'X')) == list(parameter.counts(bk, 'X'))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_self_play_learning_iteration_single_gpu():
    L = learner.ProgressPositionMovesLearn()
    L.configure({})
    po, rows = L.self_play_iteration(4096, seed=3, iteration=0, random_plies=10, device=DEV)
    params = L.read_parameters()
    assert len(params) == 37 and params[0] == 2
    assert all(-127 <= v <= 127 for v in params[1:])
    for s in range(4):
        if L.last_fits[s]['n'] > 0:
            assert max(abs(v) for v in params[1 + 9 * s: 10 + 9 * s]) in (126, 127)
    assert sum(f['n'] for f in L.last_fits) == 2 * (po.total_positions() + 4096)
    mse, score, param, nsample = L.fit_parameter(33, 48)
    assert nsample == L.last_fits[2]['n'] and len(param) == 9 and 0 <= score <= 1
    # the refit weights drive the next iteration's greedy self-play
    po2, _ = L.self_play_iteration(1024, seed=3, iteration=1, random_plies=10, device=DEV)
    assert int(po2.nplies.min()) > 0


def test_learning_is_invariant_to_how_games_are_sharded():
    """ranks emulated on one GPU: the statistics of 4 shards summed == one rank playing all games"""
    w = torch.from_numpy(parameter.ProgressPositionMovesParameter().weights_table()).to(DEV)
    whole = ops.learn_accumulate(ops.playout(8192, seed=9, gid0=0, device=DEV, policy=ops.POLICY_GREEDY,
                                             random_plies=10, weights=w))
    parts = torch.zeros_like(whole)
    for r in range(4):
        lo, hi = learner.shard_of_games(8192, r, 4)
        ops.learn_accumulate(ops.playout(hi - lo, seed=9, gid0=lo, device=DEV, policy=ops.POLICY_GREEDY,
                                         random_plies=10, weights=w), acc=parts)
    assert torch.equal(whole, parts)                              # exact integer accumulators
    a, b = learner.fit_from_stats(ops.learn_stats(whole)), learner.fit_from_stats(ops.learn_stats(parts))
    for s in range(4):
        assert np.array_equal(a[s]['coef'], b[s]['coef'])
