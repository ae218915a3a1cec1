"""Constructed line positions: for every direction, every start square and every run length, a run of
opponent discs closed by an own disc, left open, or running into the edge (the cases hands_for_direc
distinguishes, board.py:124-139), with noise discs elsewhere.  Shared by the CPU and GPU tests."""
import numpy as np

DIRS = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]     # board.py:9-17


def build():
    own, opp, sq = [], [], []
    rng = np.random.RandomState(77)
    for dx, dy in DIRS:
        for s in range(64):
            x0, y0 = s & 7, s >> 3
            for run in range(0, 8):
                for closing in ("own", "empty", "edge_or_opp"):
                    o = p = 0
                    x, y = x0 + dx, y0 + dy
                    k = 0
                    while k < run and 0 <= x < 8 and 0 <= y < 8:
                        p |= 1 << (x + 8 * y)
                        x, y, k = x + dx, y + dy, k + 1
                    if k < run:
                        continue                                  # the run does not fit on the board
                    if 0 <= x < 8 and 0 <= y < 8:
                        if closing == "own":
                            o |= 1 << (x + 8 * y)
                        elif closing == "edge_or_opp":
                            p |= 1 << (x + 8 * y)                 # one more opponent disc: run + 1, then whatever follows
                    # noise away from the line and from the move square
                    line = 0
                    lx, ly = x0, y0
                    while 0 <= lx < 8 and 0 <= ly < 8:
                        line |= 1 << (lx + 8 * ly)
                        lx, ly = lx + dx, ly + dy
                    noise_o = int(rng.randint(0, 2 ** 32)) | (int(rng.randint(0, 2 ** 32)) << 32)
                    noise_p = int(rng.randint(0, 2 ** 32)) | (int(rng.randint(0, 2 ** 32)) << 32)
                    if rng.rand() < 0.5:
                        noise_o = noise_p = 0
                    noise_o &= ~line & ~noise_p
                    noise_p &= ~line
                    own.append(o | noise_o)
                    opp.append(p | noise_p)
                    sq.append(s)
    return (np.array(own, dtype=np.uint64), np.array(opp, dtype=np.uint64), np.array(sq, dtype=np.uint8))
