"""The engine front-end speaks the line protocol game_runner.Player reads (game_runner.py:19-73).
The regexes below are the reference's; the backend here is a stub so the test needs no GPU."""
import io
import re

from subproc_b200.edax_engine import EdaxFrontEnd

GO_RE = r">(.+) plays [WB]?([a-zA-Z][0-9]|PS)"          # game_runner.py:27
PLAY_RE = r"(.+) play ([a-zA-Z][0-9]|PS|ps)"            # game_runner.py:51


class StubBackend(object):
    def __init__(self, moves):
        self.moves, self.played, self.resets = list(moves), [], 0

    def reset(self):
        self.resets += 1

    def best_move(self):
        return self.moves.pop(0)

    def play(self, hand):
        self.played.append(hand)
        return 1

    def parameter_dump(self):
        return "[1, 2, 3]"


def drive(front, out, cmd, nlines):
    start = len(out.getvalue())
    alive = front.handle(cmd + "\n")
    text = out.getvalue()[start:]
    assert text.count("\n") == nlines, (cmd, text)
    return alive, text


def test_protocol_lines_and_regexes():
    out = io.StringIO()
    be = StubBackend(['D3', 'PS'])
    fe = EdaxFrontEnd(be, name='b200', out=out)
    alive, text = drive(fe, out, 'init', 1)                       # Player.init reads 1 line (:35-39)
    assert alive and be.resets == 1
    alive, text = drive(fe, out, 'go', 3)                         # Player.go reads 3 lines (:19-33)
    m = re.findall(GO_RE, re.sub(r'[\r\n]+', "", text).rstrip())
    assert m[0] == ('b200', 'D3') and be.played == ['D3']
    alive, text = drive(fe, out, 'go', 3)
    assert re.findall(GO_RE, re.sub(r'[\r\n]+', "", text).rstrip())[0][1] == 'PS' and be.played[-1] == 'ps'
    for hand in ('c5', 'ps', 'PS'):
        alive, text = drive(fe, out, hand, 3)                     # Player.play reads 3 lines (:41-54)
        assert re.findall(PLAY_RE, text.rstrip())[0][1] == hand and be.played[-1] == hand
    alive, text = drive(fe, out, 'verbose p', 1)                  # show_hamlet_param reads 1 line (:66-73)
    assert text.strip() == "[1, 2, 3]"
    alive, text = drive(fe, out, 'verbose 0', 0)
    alive, text = drive(fe, out, 'verbose 1', 13)                 # Player.show reads 13 lines (:75-88)
    alive, text = drive(fe, out, 'quit', 1)                       # end_process reads 1 line (:56-64)
    assert not alive
