"""Live differential check: oracle vs the reference's own board.py on fresh random positions.

Runs only where the reference is reachable (/root/reference in the build container, or
oracle/_ref on a box that received it); elsewhere the golden vectors are the pin.
"""
import numpy as np
import pytest

from oracle import refshim
from oracle import make_golden as mg

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference not present on this box")


def test_random_games_against_live_reference(oracle):
    ns = refshim.load()
    rows = ns.ppml.ProgressPositionMovesParameter().default_value()
    for gid in range(100, 112):
        g = mg.play_game(ns, 11, gid, 0, 0, 0, 0, rows, record_features=True)
        pos = g['positions']
        b = np.array([int(p['b'], 16) for p in pos], dtype=np.uint64)
        w = np.array([int(p['w'], 16) for p in pos], dtype=np.uint64)
        assert [int(v) for v in oracle.puttables(b, w, 1)] == [int(p['legal_b'], 16) for p in pos]
        assert [int(v) for v in oracle.puttables(b, w, 2)] == [int(p['legal_w'], 16) for p in pos]
        assert oracle.features(b, w, 1).tolist() == [p['feat_O'] for p in pos]
        assert oracle.features(b, w, 2).tolist() == [p['feat_X'] for p in pos]
        r = oracle.playout(11, gid, 1)
        assert r['move'][:len(g['plies']), 0].tolist() == [p['move'] for p in g['plies']]


def test_greedy_games_against_live_reference(oracle):
    ns = refshim.load()
    rows = ns.ppml.ProgressPositionMovesParameter().default_value()
    for gid in range(3):
        g = mg.play_game(ns, 12, gid, 1, 6, 0, 0, rows, record_features=False)
        r = oracle.playout(12, gid, 1, policy=1, random_plies=6)
        assert r['move'][:len(g['plies']), 0].tolist() == [p['move'] for p in g['plies']]
        assert int(r['nplies'][0]) == len(g['plies'])


def test_eval_against_counts_dot(oracle):
    ns = refshim.load()
    rb = ns.board
    rng = np.random.RandomState(3)
    w = np.concatenate([rng.uniform(-2, 2, size=(4, 9)), rng.uniform(-1, 1, size=(4, 1))], axis=1)
    r = oracle.playout(13, 0, 4)
    for g in range(4):
        for t in range(0, int(r['nplies'][g]) + 1, 7):
            bb, ww = int(r['black'][t, g]), int(r['white'][t, g])
            B = rb.Board()
            for s in range(64):
                B.set(rb.Black if (bb >> s) & 1 else rb.White if (ww >> s) & 1 else rb.Empty, s & 7, s >> 3)
            for side, colour in (('O', 1), ('X', 2)):
                f = ns.ppml.counts(mg.book_of(B), side)
                row = 0 if f[0] <= 16 else 1 if f[0] <= 32 else 2 if f[0] <= 48 else 3
                want = float(np.dot(w[row, :9], np.array(f[1:], dtype=np.float64)) + w[row, 9])
                got = float(oracle.evaluate([bb], [ww], colour, w)[0])
                assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
