"""Drop-in for the reference's ``board`` module (board.py) backed by the CUDA kernels.

Same module-level names (``Empty, Black, White, DIRECS, Board, is_within_board, clone_board``) and
the same ``Board`` duck type the reference's callers rely on (game_runner.py:137-196,
game_recorder.py:64,109-112, parameter.py:5-8, parameter_progress_position_moves_learn.py:6-17,
learn_base.py:70-73): attributes ``board``, ``turn``, ``nturn``; methods ``puttables``,
``n_puttable_for``, ``is_game_over``, ``put``, ``put_s``, ``n_black/n_white/n_empty``,
``mask_count``, ``get/set``, ``hostile``, ``(de)serialize*``, ``__str__`` with the reference's
return codes (put -> 0, put_s -> -1, no exceptions for illegal moves).

Rules questions (legal moves, flips, game over, counts, counts() features) are answered by the
sm_100a kernels: every change of the position is ONE call of ``othello_board_apply_host`` -- one
launch, one stream synchronise, results written by the kernel straight into pinned host memory --
which applies the move (if any) and returns everything the reference's callers ask about the new
position before the next ply (game_runner.py:154-163 asks ~6 questions per ply).  The object keeps a
host mirror of the two bitboards so that the pure string / accessor methods need no launch.  For
throughput use ``subproc_b200.batched`` -- this class exists so the reference's single-game code
keeps working unchanged.  Without a CUDA device the rules methods raise; there is no CPU
implementation of the rules in this package.
"""
import ctypes
import re

from . import _lib

COLORS = (Empty, Black, White) = range(0, 3)          # board.py:3-7

DIRECS = (LU, U, RU, L, R, LD, D, RD) = [               # board.py:9-17
    (-1, -1), (0, -1), (1, -1),
    (-1, 0), (1, 0),
    (-1, 1), (0, 1), (1, 1)
]

_HAND_RE = re.compile(r"[WB]*([a-zA-Z])([0-9])")      # the hand-string grammar of board.py:177
_FULL = 0xFFFFFFFFFFFFFFFF


def is_within_board(x, y):                            # board.py:265-266
    return 0 <= x < 8 and 0 <= y < 8


def clone_board(board):                               # board.py:269-276
    return [[board[i][j] for j in range(8)] for i in range(8)]


def _bit(x, y):
    if not (0 <= x < 8 and 0 <= y < 8):
        raise IndexError("list index out of range")   # what board[y][x] raises in the reference
    return 1 << (x + 8 * y)


class Board(object):
    def __init__(self, device=None):
        self._device = device
        self._black = 0x0000000810000000              # board.py:25: e4, d5
        self._white = 0x0000001008000000              # board.py:24: d4, e5
        self.turn = Black
        self.nturn = 0

    # ---- device plumbing ---------------------------------------------------------------
    _contexts = {}                                     # device index -> othello_ctx (shared by all boards)
    _QUERY = 255                                       # "no move": only describe the position
    _CLASS_MASKS = (0x8100000000000081, 0x4281000000008142, 0x0042000000004200, 0x2400810000810024,
                    0x1800008181000018, 0x003C424242423C00, 0x0000240000240000, 0x0000183C3C180000)

    def _ctx_index(self):
        dev = self._device
        if dev is None:
            return 0
        if isinstance(dev, int):
            return dev
        text = str(dev)
        return int(text.split(":")[1]) if ":" in text else 0

    def _ctx(self):
        index = self._ctx_index()
        ctx = Board._contexts.get(index)
        if ctx is None:
            ctx = ctypes.c_void_p()
            rc = _lib.lib().othello_ctx_create(index, ctypes.byref(ctx))
            if rc != 0:
                raise RuntimeError("subproc_b200.board.Board needs a CUDA device: the rules run in sm_100a kernels "
                                   "and there is no CPU fallback (othello_ctx_create: %s)"
                                   % _lib.lib().othello_error_string(rc).decode())
            Board._contexts[index] = ctx
        return ctx

    def _apply(self, black, white, color, move):
        """one launch: Board.put(color, move) on (black, white) + the description of the resulting position"""
        info = _lib.PositionInfo()
        _lib.check(_lib.lib().othello_board_apply_host(self._ctx(), black, white, color, move, ctypes.byref(info)),
                   "othello_board_apply_host")
        return info

    def _info(self):
        """description of the current position (legal masks and features of both colours, counts); remembered
        until the position changes -- play_a_turn asks puttables / is_game_over / n_black for the same
        position several times (game_runner.py:137,162,194-196; game_recorder.py:112)"""
        key = (self._black, self._white)
        if getattr(self, '_info_key', None) != key:
            self._info_val = self._apply(self._black, self._white, Black, Board._QUERY)
            self._info_key = key
        return self._info_val

    def _pair(self, piece):
        """(own, opp) bitboards for colour ``piece``; hostile(piece) is Black for anything but Black."""
        return (self._black, self._white) if piece == Black else (self._white, self._black)

    def _legal_mask(self, piece):
        info = self._info()
        return info.legal_black if piece == Black else info.legal_white

    # ---- the 8x8 list view -------------------------------------------------------------
    @property
    def board(self):
        return [[self.get(x, y) for x in range(8)] for y in range(8)]

    @board.setter
    def board(self, rows):
        self._black = self._white = 0
        for y in range(8):
            for x in range(8):
                self.set(rows[y][x], x, y)

    def set(self, piece, x, y):                        # board.py:60-61
        b = _bit(x, y)
        self._black &= ~b
        self._white &= ~b
        if piece == Black:
            self._black |= b
        elif piece == White:
            self._white |= b

    def get(self, x, y):                               # board.py:63-64
        b = _bit(x, y)
        return Black if self._black & b else (White if self._white & b else Empty)

    # ---- counts ------------------------------------------------------------------------
    def _counts(self):
        info = self._info()
        return info.n_black, info.n_white, info.n_empty

    def count_over_board(self, fun):                   # board.py:29-35
        return sum(1 for y in range(8) for x in range(8) if fun(self.get(x, y)))

    def n_black(self):                                 # board.py:37-38
        return self._counts()[0]

    def n_white(self):                                 # board.py:40-41
        return self._counts()[1]

    def n_empty(self):                                 # board.py:43-44
        return self._counts()[2]

    def mask_count(self, color, mask):                 # board.py:74-81
        mask &= _FULL
        if mask in Board._CLASS_MASKS and color in (Black, White):
            # the eight square classes of counts() (parameter_progress_position_moves_learn.py:9-16) come
            # with the position description
            info = self._info()
            return (info.features_black if color == Black else info.features_white)[2 + Board._CLASS_MASKS.index(mask)]
        import torch
        from . import ops
        dev = torch.device("cuda", self._ctx_index())
        out = ops.mask_count(ops.bits_tensor([self._black], dev), ops.bits_tensor([self._white], dev),
                             torch.tensor([color], dtype=torch.uint8, device=dev), ops.bits_tensor([mask & _FULL], dev))
        return int(out.cpu()[0])

    # ---- rules -------------------------------------------------------------------------
    def puttables(self, piece):                        # board.py:46-52 (ascending x + 8*y)
        m = self._legal_mask(piece)
        return [(s & 7, s >> 3) for s in range(64) if (m >> s) & 1]

    def n_puttable_for(self, piece):                   # board.py:54-55
        return bin(self._legal_mask(piece)).count("1")

    def is_game_over(self):                            # board.py:57-58
        return self.n_puttable_for(Black) == 0 and self.n_puttable_for(White) == 0

    def is_puttable_at(self, piece, x, y):             # board.py:141-149
        return bool(self._legal_mask(piece) & _bit(x, y))

    def hostile(self, piece):                          # board.py:155-159
        return White if piece == Black else Black

    def hands_for_direc(self, direc, piece, x, y):     # board.py:124-139
        """the run put() would flip from (x, y) along ``direc``, as (piece, x, y) triples."""
        here = _bit(x, y)
        # (the reference walks the ray whatever stands on (x, y) itself)
        f = self._apply(self._black & ~here, self._white & ~here, Black if piece == Black else White, x + 8 * y).flips
        ret = []
        for i in range(1, 9):
            nx, ny = x + i * direc[0], y + i * direc[1]
            if is_within_board(nx, ny) and (f >> (nx + 8 * ny)) & 1:
                ret.append((piece, nx, ny))
            else:
                break
        return ret

    def set_hands(self, hands):                        # board.py:151-153
        for (piece, x, y) in hands:
            self.set(piece, x, y)

    def put(self, piece, x, y):                        # board.py:161-174
        _bit(x, y)                                     # IndexError like board[y][x]
        info = self._apply(self._black, self._white, Black if piece == Black else White, x + 8 * y)
        if info.ret:
            self._black, self._white = info.black, info.white
            self._info_key, self._info_val = (info.black, info.white), info      # already describes the new position
        return info.ret

    def coord_from_handstr(self, handstr):             # board.py:176-185
        b = _HAND_RE.findall(handstr)
        if len(b) > 0:
            return ord(b[0][0].lower()) - ord('a'), ord(b[0][1]) - ord('1')
        return -1, -1

    def handstr_from_coord(self, x, y):                # board.py:187-190
        return chr(ord('a') + x) + chr(ord('1') + y)

    def put_s(self, stri):                             # board.py:192-209
        out = -1
        if stri == 'PS' or stri == 'ps':
            out = 0
        else:
            x, y = self.coord_from_handstr(stri)
            if x >= 0 and y >= 0:
                out = self.put(self.turn, x, y)        # IndexError beyond h / rank 9, like board[y][x]
                if out == 0:
                    out = -1
        if out >= 0:
            self.nturn += 1                            # nturn may be a str after deserialize -> TypeError, as in the reference
            self.turn = White if self.turn == Black else Black
        return out

    # ---- text forms ----------------------------------------------------------------------
    def str_from_turn(self, color):                    # board.py:66-72
        return 'Black' if color == Black else ('White' if color == White else 'None')

    def __str__(self):                                 # board.py:94-122
        nb, nw, _ = self._counts()
        pudding = "      "
        ret = '  A B C D E F G H\n'
        for i in range(1, 9):
            ret += str(i)
            for x in range(8):
                cell = self.get(x, i - 1)
                ret += ' ' + ('*' if cell == Black else 'O' if cell == White else '.')
            if i == 4:
                ret += pudding + self.str_from_turn(self.turn) + '\'s turn'
            elif i == 5:
                ret += pudding + 'Black: ' + str(nb)
            elif i == 6:
                ret += pudding + 'White: ' + str(nw)
            ret += '\n'
        return ret

    def serialize_tuple(self):                         # board.py:211-212
        return self.board, self.turn

    def serialize_str(self, append_turn=True):         # board.py:214-221
        ret = self.serialize_board()
        if append_turn:
            ret += ' ' + self.serialize_turn()
        return ret

    def serialize_board(self):                         # board.py:223-232
        return ''.join(self.string_from_turn(self.get(s & 7, s >> 3)) for s in range(64))

    def serialize_turn(self):                          # board.py:234-235
        return self.string_from_turn(self.turn)

    def string_from_turn(self, turn):                  # board.py:237-243
        return 'O' if turn == Black else ('X' if turn == White else '-')

    def turn_from_string(self, turn_string):           # board.py:245-251
        return Black if turn_string == 'O' else (White if turn_string == 'X' else Empty)

    def deserialize(self, board_str, turn_str, nturn):   # board.py:253-262
        for i, s in enumerate(board_str):
            self.set(self.turn_from_string(s), i % 8, i // 8)
        self.turn = self.turn_from_string(turn_str)
        self.nturn = nturn

    @classmethod
    def show_mask(cls, mask):                          # board.py:83-92
        q = cls()
        for s in range(64):
            q.set(Black if (mask >> s) & 1 else Empty, s & 7, s >> 3)
        print(q)
