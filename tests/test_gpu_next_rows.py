"""SURVEY 8(f) "next" rows on a B200: book export / import in the recorder schema, the Edax-protocol
engine process driven the way game_runner.Player drives it."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from subproc_b200 import ops, books, board
from gpu_util import DEV, h

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_books_match_the_reference_recorder_strings(golden_games):
    g = golden_games[3]
    po = ops.playout(1, seed=g['seed'], gid0=g['gid'], device=DEV)
    (recs, meta), = books.books_from_playout(po)
    assert len(recs) == len(g['positions'])
    for r, p in zip(recs, g['positions']):
        assert r['book'] == p['ser'][:64] and r['whosturn'] == p['ser'][65]      # serialize_board / serialize_turn
        assert r['turn'] == p['nturn'] and r['end'] == p['over']                 # game_recorder.py:108-113
    text = books.flatfile_text(recs, meta)
    assert text.splitlines()[0] == "% Black: b200" and text.splitlines()[2] == g['positions'][0]['ser']
    b, w, who, turns = books.positions_from_books(recs, device=DEV)
    assert ops.bits_numpy(b).tolist() == [h(p['b']) for p in g['positions']]
    assert ops.bits_numpy(w).tolist() == [h(p['w']) for p in g['positions']]
    assert who.cpu().tolist() == [p['turn'] for p in g['positions']]


def test_serialize_round_trip_large():
    po = ops.playout(20000, seed=12, gid0=0, device=DEV)
    live = po.nplies >= 37                                # rows past a game's end are unspecified
    b, w = po.black[37][live].contiguous(), po.white[37][live].contiguous()
    chars = ops.serialize_boards(b, w)
    b2, w2 = ops.deserialize_boards(chars)
    assert torch.equal(b, b2) and torch.equal(w, w2)
    s = chars[5].cpu().numpy().tobytes().decode()
    bb, ww = ops.unsigned64(b[5].item()), ops.unsigned64(w[5].item())
    assert s == ''.join('O' if (bb >> i) & 1 else 'X' if (ww >> i) & 1 else '-' for i in range(64))


class Player(object):
    """the reference's Player (game_runner.py:9-64), restated for Python 3 pipes"""

    def __init__(self, cmd):
        self.proc = subprocess.Popen(cmd, shell=True, stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, cwd=ROOT)
        self.name = ''

    def _read(self, n):
        return ''.join(self.proc.stdout.readline() for _ in range(n))

    def init(self):
        self.proc.stdin.write('init\n'); self.proc.stdin.flush()
        self._read(1)

    def go(self):
        self.proc.stdin.write('go\n'); self.proc.stdin.flush()
        out = re.sub(r'[\r\n]+', "", self._read(3))
        b = re.findall(r">(.+) plays [WB]?([a-zA-Z][0-9]|PS)", out.rstrip())
        self.name = b[0][0]
        return b[0][1]

    def play(self, hand):
        self.proc.stdin.write(hand + '\n'); self.proc.stdin.flush()
        return re.findall(r"(.+) play ([a-zA-Z][0-9]|PS|ps)", self._read(3).rstrip())[0][1]

    def end_process(self):
        self.proc.stdin.write('quit\n'); self.proc.stdin.flush()
        self._read(1)
        self.proc.communicate()


def test_engine_processes_play_a_full_game_like_game_runner():
    py = sys.executable
    black = Player("%s -m subproc_b200.edax_engine --policy greedy --name greedyB200" % py)
    white = Player("%s -m subproc_b200.edax_engine --policy random --name randomB200 --seed 5" % py)
    black.init(); white.init()
    ref = board.Board()
    plies = 0
    over = ref.is_game_over()
    while not over:                                          # play_a_game / play_a_turn (game_runner.py:154-184)
        attacker, defender = (black, white) if ref.turn == board.Black else (white, black)
        ha = attacker.go().lower()
        legal = ref.puttables(ref.turn)
        if ha == 'ps':
            assert legal == []
        else:
            assert ref.coord_from_handstr(ha) in legal
        assert ref.put_s(ha) >= 0
        assert defender.play(ha) == ha
        plies += 1
        over = ref.is_game_over()
    assert black.name == 'greedyB200' and white.name == 'randomB200'
    assert plies >= 9 and ref.n_black() + ref.n_white() + ref.n_empty() == 64
    black.end_process(); white.end_process()
    assert black.proc.returncode == 0 and white.proc.returncode == 0


def test_feature_dump_cli(kat):
    out = subprocess.check_output([sys.executable, "-m", "subproc_b200.edax_engine", "-h", kat['start_serialize_str']],
                                  cwd=ROOT, text=True)
    from ast import literal_eval
    assert literal_eval(out.strip()) == kat['start_counts_O'][1:]          # parameter_learn_from_edax_protocol.py:12-13


def test_do_match_with_plugin_recorders(tmp_path, golden_games):
    """subproc.do_match (subproc.py:15-39): config-selected recorder plugin, one game, winner tuple"""
    from subproc_b200 import match, recorder
    conf = {'proc_a_path': 'random', 'proc_b_path': 'random', 'proc_n_rand_hands_for_a': 0,
            'game_recorder_from': 'subproc_b200.recorder', 'game_recorder_class': 'MemoryRecorder'}
    mem = recorder.MemoryRecorder()
    won = match.do_match(conf, seed=0, recorder=mem)
    assert won[0] in ('Black', 'White', 'None')
    lines, meta = mem.stored[0]
    g = golden_games[0]                                  # seed 0, game id 0, random engines: the same game
    assert meta['proc_a'] == 'b200-random'
    assert [r['book'] + ' ' + r['whosturn'] for r in lines] == [p['ser'] for p in g['positions']]
    conf.update(proc_a_path='greedy', proc_b_path='greedy', proc_n_rand_hands_for_a=3, proc_n_rand_hands_for_b=3)
    rec = match.get_game_recorder(conf)
    assert isinstance(rec, recorder.MemoryRecorder)
    won = match.do_match(conf, seed=5)
    assert won[0] in ('Black', 'White', 'None')


def test_features_through_the_engine_process_parameter_class(golden_games):
    """LearnFromEdaxProtocolProcessParameter shells out to `<engine> -h "<sfen>"` like the reference
    (parameter_learn_from_edax_protocol.py:7-13,35-39); our engine answers with the GPU features"""
    from subproc_b200 import parameter
    P = parameter.LearnFromEdaxProtocolProcessParameter()
    P.configure({'learn_learn_for_path': "cd %s && %s -m subproc_b200.edax_engine" % (ROOT, sys.executable)})
    assert P.header() == 3
    p = golden_games[0]['positions'][17]
    bk = {'book': p['ser'][:64], 'whosturn': p['ser'][65], 'turn': p['nturn']}
    want = p['feat_O'] if p['ser'][65] == 'O' else p['feat_X']
    assert P.hash_from_book(bk, 'O') == ':'.join(str(v) for v in want)
