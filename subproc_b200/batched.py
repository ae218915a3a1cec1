"""BatchedOthello: the reference's ``Board`` (board.py:20-263) for B games at once, state in HBM.

State is SoA: ``black``/``white`` int64[B] bit patterns (bit s = x + 8*y), ``turn`` uint8[B]
(1 Black / 2 White), ``nturn`` int32[B].  Every rules method is one kernel launch over the batch.
Method names follow Board so that code written against the single-game object reads the same.
"""
import torch

from . import ops
from .ops import BLACK, WHITE, PASS


class BatchedOthello(object):
    def __init__(self, n, device=None, black=None, white=None, turn=None, nturn=None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedOthello needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.n = int(n)
        full = lambda v, dt: torch.full((self.n,), v, dtype=dt, device=self.device)
        # Board.__init__ (board.py:22-27): standard opening, Black to move, nturn = 0
        self.black = full(ops.signed64(ops.START_BLACK), torch.int64) if black is None else black.to(self.device).contiguous()
        self.white = full(ops.signed64(ops.START_WHITE), torch.int64) if white is None else white.to(self.device).contiguous()
        self.turn = full(BLACK, torch.uint8) if turn is None else turn.to(self.device).contiguous()
        self.nturn = full(0, torch.int32) if nturn is None else nturn.to(self.device).contiguous()
        self.flags = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        self._flips = torch.empty(self.n, dtype=torch.int64, device=self.device)
        self._ret = torch.empty(self.n, dtype=torch.int32, device=self.device)

    # ---- rules ---------------------------------------------------------------------------
    def _pair(self, piece):
        return (self.black, self.white) if piece == BLACK else (self.white, self.black)

    def puttables(self, piece):
        """legal-move masks of colour ``piece`` for every game (board.py:46-52)."""
        own, opp = self._pair(piece)
        return ops.legal(own, opp)

    def puttables_for_turn(self):
        black_to_move = self.turn == BLACK
        own = torch.where(black_to_move, self.black, self.white)
        opp = torch.where(black_to_move, self.white, self.black)
        return ops.legal(own, opp)

    def n_puttable_for(self, piece):
        """mobility (board.py:54-55) as int32[B]."""
        m = self.puttables(piece)
        return ops.counts(m, torch.zeros_like(m))[:, 0].contiguous()

    def is_game_over(self):
        """bool[B] (board.py:57-58)."""
        return (self.puttables(BLACK) == 0) & (self.puttables(WHITE) == 0)

    def put_s(self, move):
        """Board.put_s with decoded hands (uint8[B]: square 0..63, 64 = pass): in place.

        Returns int32[B]: -1 illegal (that game untouched), 0 pass, else discs flipped
        (board.py:192-209).  ``self.flags`` then holds OTHELLO_F_MUST_PASS / OTHELLO_F_GAME_OVER.
        """
        ops.step(self.black, self.white, self.turn, self.nturn, move, self._flips, self._ret, self.flags)
        return self._ret

    def last_flips(self):
        return self._flips

    # ---- counts / features -----------------------------------------------------------------
    def counts(self):
        return ops.counts(self.black, self.white)

    def n_black(self):
        return self.counts()[:, 0]

    def n_white(self):
        return self.counts()[:, 1]

    def n_empty(self):
        return self.counts()[:, 2]

    def features(self, side):
        """counts(a_book, side) for every game: int32[B][10]."""
        s = torch.full((self.n,), side, dtype=torch.uint8, device=self.device)
        return ops.features(self.black, self.white, s)

    def evaluate(self, side, weights):
        s = torch.full((self.n,), side, dtype=torch.uint8, device=self.device)
        return ops.evaluate(self.black, self.white, s, weights)

    # ---- text forms (host side) --------------------------------------------------------------
    def serialize_board(self, i):
        """64-char 'O'/'X'/'-' string of game i (board.py:223-243)."""
        b = ops.unsigned64(self.black[i].item())
        w = ops.unsigned64(self.white[i].item())
        return ''.join('O' if (b >> s) & 1 else ('X' if (w >> s) & 1 else '-') for s in range(64))
