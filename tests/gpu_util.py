"""helpers shared by the -m gpu tests."""
import numpy as np
import torch

from subproc_b200 import ops

DEV = "cuda:0"


def dev_bits(a):
    return ops.bits_tensor(np.asarray(a, dtype=np.uint64), DEV)


def dev_u8(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.uint8))).to(DEV)


def dev_i32(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int32))).to(DEV)


def host_bits(t):
    return ops.bits_numpy(t)


def h(s):
    return int(s, 16)


def sample_positions(oracle, n_games, seed, stride=1):
    """positions (black, white, turn) drawn from oracle random playouts: reachable, all phases."""
    r = oracle.playout(seed, 0, n_games)
    bs, ws, ts = [], [], []
    for g in range(n_games):
        n = int(r['nplies'][g])
        idx = np.arange(0, n + 1, stride)
        bs.append(r['black'][idx, g])
        ws.append(r['white'][idx, g])
    b = np.concatenate(bs)
    w = np.concatenate(ws)
    return b, w
