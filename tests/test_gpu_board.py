"""The single-game Board facade (drop-in for the reference's board module) on a B200."""
import pytest

from subproc_b200 import board
from gpu_util import h

pytestmark = pytest.mark.gpu


def test_kat_start_and_put_s_cases(kat):
    b = board.Board()
    assert [list(c) for c in b.puttables(board.Black)] == kat['start_puttables_black']
    assert [list(c) for c in b.puttables(board.White)] == kat['start_puttables_white']
    assert str(b) == kat['start_str']
    assert (b.n_black(), b.n_white(), b.n_empty()) == (2, 2, 60)
    assert not b.is_game_over()
    for c in kat['put_s_cases']:
        q = board.Board()
        if c['err'] is not None:
            with pytest.raises(IndexError):
                q.put_s(c['s'])
            continue
        assert q.put_s(c['s']) == c['ret'], c['s']
        assert ("%016x" % q._black, "%016x" % q._white, q.turn, q.nturn) == (c['b'], c['w'], c['turn'], c['nturn'])
    q = board.Board()
    assert q.put_s('d3') == kat['after_d3']['ret']
    assert str(q) == kat['after_d3']['str'] and q.serialize_str() == kat['after_d3']['ser']
    assert (q.n_black(), q.n_white(), q.turn, q.nturn) == (kat['after_d3']['nb'], kat['after_d3']['nw'],
                                                          kat['after_d3']['turn'], kat['after_d3']['nturn'])


def test_replay_a_golden_game_through_the_facade(golden_games):
    g = golden_games[0]
    b = board.Board()
    for ply, pos in zip(g['plies'], g['positions'][1:]):
        legal = b.puttables(b.turn)
        assert (len(legal) == 0) == (ply['hand'] == 'ps')
        assert b.put_s(ply['hand']) == ply['ret']
        assert ("%016x" % b._black, "%016x" % b._white) == (pos['b'], pos['w'])
        assert b.serialize_str() == pos['ser'] and b.is_game_over() == pos['over']
    assert (b.n_black(), b.n_white(), b.n_empty()) == (pos['nb'], pos['nw'], pos['ne'])


def test_put_any_colour_mask_count_and_rays(golden_probe):
    rec = golden_probe[0]
    for side in rec['put']:
        for s in range(0, 64, 3):
            q = board.Board()
            q._black, q._white = h(rec['b']), h(rec['w'])
            assert q.put(side['piece'], s & 7, s >> 3) == side['ret'][s]
            own = q._black if side['piece'] == board.Black else q._white
            assert own & h(side['flips'][s]) == h(side['flips'][s])
            q._black, q._white = h(rec['b']), h(rec['w'])
            if q.get(s & 7, s >> 3) == board.Empty:
                total = sum(len(q.hands_for_direc(d, side['piece'], s & 7, s >> 3)) for d in board.DIRECS)
                assert total == side['ret'][s]
                assert q.is_puttable_at(side['piece'], s & 7, s >> 3) == (side['ret'][s] > 0)
    q = board.Board()
    q._black, q._white = h(rec['b']), h(rec['w'])
    assert [q.mask_count(board.Black, h(rec['mask'])), q.mask_count(board.White, h(rec['mask']))] == rec['mask_count']


def test_game_runner_single_game_with_recorder():
    from subproc_b200.game_runner import GameRunner

    class Rec(object):
        def __init__(self):
            self.lines, self.meta, self.stored = [], {}, False

        def add(self, b):
            self.lines.append((b.serialize_board(), b.serialize_turn(), b.nturn, b.is_game_over()))

        def add_meta(self, m):
            self.meta.update(m)

        def store(self):
            self.stored = True

    rec = Rec()
    won = GameRunner('random', 'random', rec, False, 0, 0, seed=0).play_a_game()
    assert rec.stored and rec.meta['proc_a'] == 'b200-random'
    assert rec.lines[0][0] == '---------------------------XO------OX---------------------------'
    assert [l[2] for l in rec.lines] == list(range(len(rec.lines)))
    assert rec.lines[-1][3] and not any(l[3] for l in rec.lines[:-1])
    assert won[0] in ("Black", "White", "None")
