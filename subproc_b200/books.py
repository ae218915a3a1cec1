"""Trajectories <-> the reference's "book" records (the recorder / reader schema).

A recorded game in the reference is a list of per-position dicts
``{'book': <64 chars 'O'/'X'/'-'>, 'whosturn': 'O'|'X', 'turn': nturn, 'end': bool}``
(RedisRecorder.add, game_recorder.py:107-114) plus a meta dict (game_runner.py:186-187); the flat-file
recorder writes ``serialize_str()`` lines under two '%' header lines (game_recorder.py:64-76).  The
learner reads the same dicts back (replearn.py:27-46, parameter.py:5-8).  The 64-character strings
are produced / parsed on the GPU (othello_serialize_boards / othello_deserialize_boards); only the
selected games cross PCIe.
"""
import numpy as np
import torch

from . import ops


def _turn_chars(first_turn, n):
    a, b = ('O', 'X') if first_turn == ops.BLACK else ('X', 'O')
    return [a if t % 2 == 0 else b for t in range(n)]


def books_from_playout(po, games=None, first_turn=None):
    """[(book_records, meta)] for the selected games of a Playout (default: all).

    book_records are in playing order (turn 0 first); ``learn_books`` reverses them itself
    (replearn.py:37-38)."""
    if po.black is None:
        raise ValueError("books_from_playout needs a Playout with trajectories")
    idx = torch.arange(po.n_games, device=po.nplies.device) if games is None else \
        torch.as_tensor(list(games), dtype=torch.int64, device=po.nplies.device)
    nplies = po.nplies[idx].cpu().numpy()
    t_hi = int(min(nplies.max(), po.t_max)) + 1 if len(nplies) else 0
    blacks = po.black[:t_hi][:, idx].t().contiguous().reshape(-1)          # [game][t]
    whites = po.white[:t_hi][:, idx].t().contiguous().reshape(-1)
    chars = ops.serialize_boards(blacks, whites).cpu().numpy().reshape(len(nplies), t_hi, 64)
    out = []
    for k in range(len(nplies)):
        n = int(min(nplies[k], po.t_max))
        ft = ops.BLACK if first_turn is None else int(first_turn[k])
        tc = _turn_chars(ft, n + 1)
        recs = [{'book': chars[k, t].tobytes().decode('ascii'), 'whosturn': tc[t], 'turn': t, 'end': t == n}
                for t in range(n + 1)]
        out.append((recs, {'proc_a': 'b200', 'proc_b': 'b200', 'hamletparam': 'No Hamlet'}))
    return out


def flatfile_text(records, meta):
    """the file FlatFileRecorder.store writes (game_recorder.py:67-76)"""
    lines = ["% Black: " + meta['proc_a'], "% White: " + meta['proc_b']]
    lines += [r['book'] + ' ' + r['whosturn'] for r in records]
    return "\n".join(lines) + "\n"


def positions_from_books(records, device=None):
    """book dicts -> (black, white int64 tensors, whosturn uint8 tensor, turn list): Board.deserialize
    (board.py:253-262) of every record in one launch."""
    device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
    raw = np.frombuffer("".join(r['book'] for r in records).encode('ascii'), dtype=np.uint8).reshape(len(records), 64)
    chars = torch.from_numpy(raw.copy()).to(device)
    black, white = ops.deserialize_boards(chars)
    who = torch.tensor([ops.BLACK if r['whosturn'] == 'O' else ops.WHITE if r['whosturn'] == 'X' else ops.EMPTY
                        for r in records], dtype=torch.uint8, device=device)
    return black, white, who, [r['turn'] for r in records]
