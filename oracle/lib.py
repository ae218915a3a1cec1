"""oracle/lib.py -- ctypes + numpy wrapper around oracle/othello_oracle.c (TEST INFRASTRUCTURE).

Builds ``oracle/liboracle.so`` with gcc on first use (``build()``), then exposes batch calls on
numpy arrays.  Bitboards are uint64 arrays, bit s = x + 8*y (board.py:79); colours 1 = Black,
2 = White (board.py:3-7); moves are uint8 squares 0..63, 64 = pass.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "othello_oracle.c")
SO = os.path.join(HERE, "liboracle.so")

START_BLACK = 0x0000000810000000     # board.py:25 (d5, e4)
START_WHITE = 0x0000001008000000     # board.py:24 (d4, e5)
BLACK, WHITE = 1, 2
PASS = 64
POLICY_RANDOM, POLICY_GREEDY = 0, 1

# parameter_progress_position_moves_learn.py:30-36, as [4][10] with a zero intercept column
DEFAULT_WEIGHTS = np.array([
    [100, 99, -1, -1, -1, -1, 3, 8, 20, 0],
    [75, 99, 2, -5, 7, 6, 4, 5, 5, 0],
    [25, 99, 2, -5, -7, -6, 4, 5, 5, 0],
    [1, 100, 50, 30, 30, 30, 30, 30, 30, 0]], dtype=np.float64)

_lib = None


def build(force=False):
    if force or not os.path.isfile(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", SO, SRC, "-lm"])
    return SO


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        L.orc_puttables.restype = ctypes.c_uint64
        L.orc_puttables.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
        L.orc_perft.restype = ctypes.c_uint64
        L.orc_perft.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
        L.orc_rng_key.restype = ctypes.c_uint32
        L.orc_rng_key.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.orc_rng_draw.restype = ctypes.c_uint32
        L.orc_rng_draw.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.orc_target.restype = ctypes.c_double
        L.orc_target.argtypes = [ctypes.c_int, ctypes.c_int]
        L.orc_smooth.restype = ctypes.c_double
        L.orc_smooth.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_mask_count.restype = ctypes.c_int
        L.orc_mask_count.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64]
        L.orc_playout_batch.restype = None
        L.orc_playout_batch.argtypes = [
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint8),
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int64,
            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint8),
            ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        for name in ("orc_puttables_batch", "orc_game_over_batch", "orc_step_batch", "orc_put_batch",
                     "orc_counts_batch", "orc_features_batch", "orc_eval_batch"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def _u8(a, n=None):
    a = np.asarray(a, dtype=np.uint8)
    if a.ndim == 0 and n is not None:
        a = np.full(n, int(a), dtype=np.uint8)
    return np.ascontiguousarray(a)


def puttables(black, white, piece):
    """legal-move masks for colour ``piece`` (board.py:46-52)."""
    black, white = _u64(black), _u64(white)
    piece = _u8(piece, black.size)
    out = np.zeros(black.size, dtype=np.uint64)
    lib().orc_puttables_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(piece, ctypes.c_uint8),
                              _p(out, ctypes.c_uint64), ctypes.c_int64(black.size))
    return out


def game_over(black, white):
    black, white = _u64(black), _u64(white)
    out = np.zeros(black.size, dtype=np.uint8)
    lib().orc_game_over_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(out, ctypes.c_uint8),
                              ctypes.c_int64(black.size))
    return out


def step(black, white, turn, nturn, move):
    """put_s (board.py:192-209) on copies; returns (black, white, turn, nturn, flips, ret)."""
    black, white = _u64(black).copy(), _u64(white).copy()
    turn = _u8(turn, black.size).copy()
    nturn = np.ascontiguousarray(np.asarray(nturn, dtype=np.int32)).copy()
    move = _u8(move, black.size)
    flips = np.zeros(black.size, dtype=np.uint64)
    ret = np.zeros(black.size, dtype=np.int32)
    lib().orc_step_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(turn, ctypes.c_uint8),
                         _p(nturn, ctypes.c_int32), _p(move, ctypes.c_uint8), _p(flips, ctypes.c_uint64),
                         _p(ret, ctypes.c_int32), ctypes.c_int64(black.size))
    return black, white, turn, nturn, flips, ret


def put(black, white, piece, square):
    """put(piece, x, y) (board.py:161-174) on copies; returns (black, white, flips, ret)."""
    black, white = _u64(black).copy(), _u64(white).copy()
    piece = _u8(piece, black.size)
    square = _u8(square, black.size)
    flips = np.zeros(black.size, dtype=np.uint64)
    ret = np.zeros(black.size, dtype=np.int32)
    lib().orc_put_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(piece, ctypes.c_uint8),
                        _p(square, ctypes.c_uint8), _p(flips, ctypes.c_uint64), _p(ret, ctypes.c_int32),
                        ctypes.c_int64(black.size))
    return black, white, flips, ret


def counts(black, white):
    """[n][3] = n_black, n_white, n_empty (board.py:37-44)."""
    black, white = _u64(black), _u64(white)
    out = np.zeros((black.size, 3), dtype=np.int32)
    lib().orc_counts_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(out, ctypes.c_int32),
                           ctypes.c_int64(black.size))
    return out


def mask_count(black, white, color, mask):
    return lib().orc_mask_count(int(black), int(white), int(color), int(mask))


def features(black, white, side):
    """[n][10] = counts(a_book, side) (parameter_progress_position_moves_learn.py:5-17)."""
    black, white = _u64(black), _u64(white)
    side = _u8(side, black.size)
    out = np.zeros((black.size, 10), dtype=np.int32)
    lib().orc_features_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(side, ctypes.c_uint8),
                             _p(out, ctypes.c_int32), ctypes.c_int64(black.size))
    return out


def evaluate(black, white, side, weights):
    """fp64 linear form w[phase] . (mobility, a..h) + intercept; weights [4][10]."""
    black, white = _u64(black), _u64(white)
    side = _u8(side, black.size)
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(4, 10))
    out = np.zeros(black.size, dtype=np.float64)
    lib().orc_eval_batch(_p(black, ctypes.c_uint64), _p(white, ctypes.c_uint64), _p(side, ctypes.c_uint8),
                         _p(w, ctypes.c_double), _p(out, ctypes.c_double), ctypes.c_int64(black.size))
    return out


def perft(depth, black=START_BLACK, white=START_WHITE, turn=BLACK):
    return int(lib().orc_perft(black, white, turn, depth))


def rng_key(seed, gid):
    return int(lib().orc_rng_key(seed, gid))


def rng_draw(key, ply, stream):
    return int(lib().orc_rng_draw(key, ply, stream))


def playout(seed, gid0, n, black0=None, white0=None, turn0=None, policy=POLICY_RANDOM, random_plies=0,
            n_rand_black=0, n_rand_white=0, weights=None, t_max=120, trajectory=True, policy_white=None,
            weights_white=None):
    """Play games gid0..gid0+n-1 (GameRunner.play_a_game, game_runner.py:165-201).

    Returns dict(black[t_max+1][n], white[t_max+1][n], move[t_max][n], nplies[n], final_black[n],
    final_white[n]); trajectory arrays are zero beyond each game's length (move = 255).
    """
    w = np.ascontiguousarray(np.asarray(DEFAULT_WEIGHTS if weights is None else weights,
                                        dtype=np.float64).reshape(4, 10))
    ww = w if weights_white is None else np.ascontiguousarray(np.asarray(weights_white, dtype=np.float64).reshape(4, 10))
    pw = policy if policy_white is None else policy_white
    b0 = _u64(black0) if black0 is not None else None
    w0 = _u64(white0) if white0 is not None else None
    t0 = _u8(turn0, n) if turn0 is not None else None
    tb = np.zeros((t_max + 1, n), dtype=np.uint64) if trajectory else None
    tw = np.zeros((t_max + 1, n), dtype=np.uint64) if trajectory else None
    mv = np.full((t_max, n), 255, dtype=np.uint8) if trajectory else None
    nplies = np.zeros(n, dtype=np.int32)
    fb = np.zeros(n, dtype=np.uint64)
    fw = np.zeros(n, dtype=np.uint64)
    lib().orc_playout_batch(ctypes.c_uint64(seed), ctypes.c_uint64(gid0), ctypes.c_int64(n),
                            _p(b0, ctypes.c_uint64), _p(w0, ctypes.c_uint64), _p(t0, ctypes.c_uint8),
                            policy, pw, random_plies, n_rand_black, n_rand_white,
                            _p(w, ctypes.c_double), _p(ww, ctypes.c_double), t_max, ctypes.c_int64(n),
                            _p(tb, ctypes.c_uint64), _p(tw, ctypes.c_uint64), _p(mv, ctypes.c_uint8),
                            _p(nplies, ctypes.c_int32), _p(fb, ctypes.c_uint64), _p(fw, ctypes.c_uint64))
    return dict(black=tb, white=tw, move=mv, nplies=nplies, final_black=fb, final_white=fw)


def target(value, turn_left):
    return float(lib().orc_target(int(value), int(turn_left)))


def smooth(current, new_value):
    return float(lib().orc_smooth(float(current), float(new_value)))
