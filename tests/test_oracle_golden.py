"""Pins the oracle (oracle/othello_oracle.c) to vectors produced by the reference's own board.py.

The vectors in tests/golden/ were written by oracle/make_golden.py from /root/reference; this
file is what turns the oracle from "a port" into "a pinned port".
"""
import numpy as np

from oracle import make_golden as mg     # only for the pure-Python RNG restatement


def h(s):
    return int(s, 16)


def test_kat_start_position(kat, oracle):
    assert h(kat['start_black']) == oracle.START_BLACK == 0x0000000810000000
    assert h(kat['start_white']) == oracle.START_WHITE == 0x0000001008000000
    for piece, key in ((1, 'start_puttables_black'), (2, 'start_puttables_white')):
        want = 0
        for x, y in kat[key]:
            want |= 1 << (x + 8 * y)
        got = int(oracle.puttables([oracle.START_BLACK], [oracle.START_WHITE], piece)[0])
        assert got == want
    assert kat['start_puttables_black'] == [[3, 2], [2, 3], [5, 4], [4, 5]]
    assert list(oracle.features([oracle.START_BLACK], [oracle.START_WHITE], 1)[0]) == kat['start_counts_O']
    assert list(oracle.features([oracle.START_BLACK], [oracle.START_WHITE], 2)[0]) == kat['start_counts_X']
    assert kat['start_counts_O'] == [4, 4, 0, 0, 0, 0, 0, 0, 0, 2]


def test_kat_perft(kat, oracle):
    assert kat['perft'] == [4, 12, 56, 244, 1396, 8200]
    assert [oracle.perft(d) for d in range(1, 7)] == kat['perft']


def test_perft_deeper_known_answers(oracle):
    # depths 7, 8 (9, 10 are checked on the GPU against these too)
    assert oracle.perft(7) == 55092
    assert oracle.perft(8) == 390216


def test_default_weights_match_reference(kat, oracle):
    assert kat['header'] == 2
    assert np.array_equal(oracle.DEFAULT_WEIGHTS[:, :9], np.array(kat['default_value'], dtype=np.float64))


def test_put_s_cases(kat, oracle):
    for c in kat['put_s_cases']:
        if c['err'] is not None:
            continue                     # IndexError cases belong to the string facade (test_board_host.py)
        x, y = c['coord']
        if c['s'] in ('ps', 'PS'):
            move = 64
        elif x >= 0 and y >= 0:
            move = x + 8 * y
        else:
            move = 255
        b, w, t, nt, fl, ret = oracle.step([oracle.START_BLACK], [oracle.START_WHITE], 1, [0], move)
        assert int(ret[0]) == c['ret'], c
        assert (int(b[0]), int(w[0]), int(t[0]), int(nt[0])) == (h(c['b']), h(c['w']), c['turn'], c['nturn']), c


def test_games_every_ply(golden_games, oracle):
    for g in golden_games:
        pos, plies = g['positions'], g['plies']
        b = np.array([h(p['b']) for p in pos], dtype=np.uint64)
        w = np.array([h(p['w']) for p in pos], dtype=np.uint64)
        assert [hex(int(v)) for v in oracle.puttables(b, w, 1)] == [hex(h(p['legal_b'])) for p in pos]
        assert [hex(int(v)) for v in oracle.puttables(b, w, 2)] == [hex(h(p['legal_w'])) for p in pos]
        assert list(oracle.game_over(b, w)) == [int(p['over']) for p in pos]
        assert oracle.counts(b, w).tolist() == [[p['nb'], p['nw'], p['ne']] for p in pos]
        if 'feat_O' in pos[0]:
            assert oracle.features(b, w, 1).tolist() == [p['feat_O'] for p in pos]
            assert oracle.features(b, w, 2).tolist() == [p['feat_X'] for p in pos]
        n = len(plies)
        turn = np.array([p['turn'] for p in pos[:n]], dtype=np.uint8)
        nturn = np.array([p['nturn'] for p in pos[:n]], dtype=np.int32)
        move = np.array([p['move'] for p in plies], dtype=np.uint8)
        b2, w2, t2, nt2, fl, ret = oracle.step(b[:n], w[:n], turn, nturn, move)
        assert ret.tolist() == [p['ret'] for p in plies]
        assert [int(v) for v in fl] == [h(p['flips']) for p in plies]
        assert np.array_equal(b2, b[1:]) and np.array_equal(w2, w[1:])
        assert t2.tolist() == [p['turn'] for p in pos[1:]]
        assert nt2.tolist() == [p['nturn'] for p in pos[1:]]
        assert pos[-1]['over'] and not any(p['over'] for p in pos[:-1])


def test_playout_loop_reproduces_reference_games(golden_games, oracle):
    """The oracle's game loop (go_for substitution, engines, RNG) replays the reference-driven games."""
    for g in golden_games:
        ww = None
        if 'rows_white' in g:
            ww = np.concatenate([np.array(g['rows_white'], dtype=np.float64), np.zeros((4, 1))], axis=1)
        r = oracle.playout(g['seed'], g['gid'], 1, policy=g['policy'], random_plies=g['random_plies'],
                           n_rand_black=g['n_rand_black'], n_rand_white=g['n_rand_white'],
                           policy_white=g.get('policy_white'), weights_white=ww)
        n = len(g['plies'])
        assert int(r['nplies'][0]) == n
        assert r['move'][:n, 0].tolist() == [p['move'] for p in g['plies']]
        assert [int(v) for v in r['black'][:n + 1, 0]] == [h(p['b']) for p in g['positions']]
        assert [int(v) for v in r['white'][:n + 1, 0]] == [h(p['w']) for p in g['positions']]
        assert int(r['final_black'][0]) == h(g['positions'][-1]['b'])
        assert int(r['final_white'][0]) == h(g['positions'][-1]['w'])


def test_probe_put_both_colours_and_mask_count(golden_probe, oracle):
    for rec in golden_probe:
        b0, w0 = h(rec['b']), h(rec['w'])
        sq = np.arange(64, dtype=np.uint8)
        for side in rec['put']:
            b2, w2, fl, ret = oracle.put(np.full(64, b0, np.uint64), np.full(64, w0, np.uint64), side['piece'], sq)
            assert ret.tolist() == side['ret']
            assert [int(v) for v in fl] == [h(s) for s in side['flips']]
        m = h(rec['mask'])
        assert [oracle.mask_count(b0, w0, 1, m), oracle.mask_count(b0, w0, 2, m)] == rec['mask_count']
        _, _, _, _, _, ret = oracle.step(np.full(64, b0, np.uint64), np.full(64, w0, np.uint64), rec['turn'],
                                         np.zeros(64, np.int32), sq)
        assert ret.tolist() == rec['put_s']


def test_rng_spec(oracle):
    for seed, gid in ((0, 0), (1, 5), (2 ** 40 + 3, 2 ** 33 + 9), (2 ** 64 - 1, 2 ** 64 - 1)):
        key = mg.rng_key(seed, gid)
        assert oracle.rng_key(seed, gid) == key
        for ply in (0, 1, 59, 119):
            for stream in (0, 1):
                assert oracle.rng_draw(key, ply, stream) == mg.rng_draw(key, ply, stream)


def test_learner_arithmetic(oracle):
    # progress_position_moves_learn.py:55 and :56-61, evaluated by CPython itself
    for value in (-64, -3, 0, 7, 64):
        for k in (0, 1, 17, 60, 119):
            assert oracle.target(value, k) == float(value) * (0.90 ** k)
    assert oracle.smooth(0.0, 5.5) == 5.5
    cur, new = 12.25, -3.5
    assert oracle.smooth(cur, new) == cur * (1 - 0.03) + new * 0.03


def test_value_table_restatement_against_the_reference_text(oracle):
    """tests/golden/value_table.json.gz was produced by EXECUTING the reference's own
    __update_state_for_a_book / __update_state_map (progress_position_moves_learn.py:37-62, cut out of the
    file by oracle/make_golden.py::reference_update_rule) on the first 24 golden games.  The oracle's
    restatement of the target and the smoothing rule must land on the same bits."""
    from conftest import load_golden
    gold = load_golden("value_table.json.gz")
    assert (gold['a'], gold['l'], gold['seed']) == (0.03, 0.90, 0)
    table = {}
    for (lo, hi), want in zip(gold['batches'], gold['after']):
        ref = oracle.playout(gold['seed'], lo, hi - lo)
        for g in range(hi - lo):
            L = int(ref['nplies'][g])
            b, w = ref['black'][:L + 1, g], ref['white'][:L + 1, g]
            fo, fx = oracle.features(b, w, 1), oracle.features(b, w, 2)
            nb = bin(int(ref['final_black'][g])).count('1')
            nw = bin(int(ref['final_white'][g])).count('1')
            for t in range(L, -1, -1):                               # terminal record first (replearn.py:37-38)
                for feats, value in ((fo[t], nb - nw), (fx[t], nw - nb)):      # 'O' then 'X' (:44-47)
                    key = ':'.join(str(int(v)) for v in feats)
                    table[key] = oracle.smooth(table.get(key, 0.0), oracle.target(value, L - t))
        assert set(table) == set(want)
        assert all(table[k].hex() == want[k] for k in want)
