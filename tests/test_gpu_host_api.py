"""The host-buffer front end of the C ABI (othello_*_host): what a non-torch caller binds
(INTEGRATION.md).  Plain numpy buffers in, plain numpy buffers out; compared with the oracle and with
the device-pointer path."""
import ctypes

import numpy as np
import pytest
import torch

from subproc_b200 import _lib, ops
from gpu_util import DEV, sample_positions

pytestmark = pytest.mark.gpu


def P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def ctx():
    L = _lib.lib()
    c = ctypes.c_void_p()
    assert L.othello_ctx_create(0, ctypes.byref(c)) == 0
    yield c
    L.othello_ctx_destroy(c)


def test_legal_and_step_host(ctx, oracle):
    L = _lib.lib()
    b, w = sample_positions(oracle, 60, seed=61)
    n = b.size
    out = np.zeros(n, np.uint64)
    assert L.othello_legal_host(ctx, P(b), P(w), P(out), n) == 0
    assert np.array_equal(out, oracle.puttables(b, w, 1))
    rng = np.random.RandomState(2)
    turn = rng.randint(1, 3, size=n).astype(np.uint8)
    nturn = rng.randint(0, 60, size=n).astype(np.int32)
    move = rng.randint(0, 66, size=n).astype(np.uint8)
    wb, ww, wt, wnt, wfl, wret = oracle.step(b, w, turn, nturn, move)
    hb, hw, ht, hnt = b.copy(), w.copy(), turn.copy(), nturn.copy()
    fl, ret, flags = np.zeros(n, np.uint64), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    assert L.othello_step_host(ctx, P(hb), P(hw), P(ht), P(hnt), P(move), P(fl), P(ret), P(flags), n) == 0
    assert np.array_equal(hb, wb) and np.array_equal(hw, ww) and np.array_equal(ht, wt) and np.array_equal(hnt, wnt)
    assert np.array_equal(fl, wfl) and np.array_equal(ret, wret)
    assert L.othello_legal_host(ctx, None, None, None, 0) == 0 and L.othello_legal_host(ctx, None, None, None, 3) == -1


def test_playout_host_small_with_host_trajectory(ctx, oracle):
    L = _lib.lib()
    n, t_max = 777, 70
    b0, w0 = sample_positions(oracle, 30, seed=62, stride=2)
    b0, w0 = np.ascontiguousarray(b0[:n]), np.ascontiguousarray(w0[:n])
    turn0 = (np.arange(n) % 2 + 1).astype(np.uint8)
    tb, tw = np.zeros((t_max + 1, n), np.uint64), np.zeros((t_max + 1, n), np.uint64)
    tm = np.zeros((t_max, n), np.uint8)
    npl, fb, fw = np.zeros(n, np.int32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    wts = oracle.DEFAULT_WEIGHTS.astype(np.float32)
    assert L.othello_playout_host(ctx, 7, 1000, n, P(b0), P(w0), P(turn0), 1, 4, 2, 3, P(wts), -1, None, t_max,
                                  P(tb), P(tw), P(tm), P(npl), P(fb), P(fw)) == 0
    ref = oracle.playout(7, 1000, n, black0=b0, white0=w0, turn0=turn0, policy=1, random_plies=4, n_rand_black=2,
                         n_rand_white=3, t_max=t_max)
    assert np.array_equal(npl, ref['nplies']) and np.array_equal(fb, ref['final_black']) and np.array_equal(fw, ref['final_white'])
    t_idx = np.arange(t_max + 1)[:, None]
    vp = t_idx <= np.minimum(npl, t_max)[None, :]
    vm = t_idx[:-1] < np.minimum(npl, t_max)[None, :]
    assert np.array_equal(tb[vp], ref['black'][vp]) and np.array_equal(tw[vp], ref['white'][vp])
    assert np.array_equal(tm[vm], ref['move'][vm])


def test_playout_host_chunked_pipeline_equals_one_launch(ctx):
    """n large enough for several chunks on rotating streams; results must equal the single-launch path"""
    L = _lib.lib()
    n = 300000 + 17
    npl, fb, fw = np.zeros(n, np.int32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    assert L.othello_playout_host(ctx, 3, 55, n, None, None, None, 0, 0, 0, 0, None, -1, None, 120, None, None, None,
                                  P(npl), P(fb), P(fw)) == 0
    po = ops.playout(n, seed=3, gid0=55, device=DEV, trajectory=False)
    assert np.array_equal(npl, po.nplies.cpu().numpy())
    assert np.array_equal(fb, ops.bits_numpy(po.final_black)) and np.array_equal(fw, ops.bits_numpy(po.final_white))
    # the trajectory of the last host call stays in device memory owned by the context
    tbp, twp, tmp = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    stride, tmax = ctypes.c_int64(), ctypes.c_int32()
    assert L.othello_ctx_trajectory(ctx, ctypes.byref(tbp), ctypes.byref(twp), ctypes.byref(tmp), ctypes.byref(stride),
                                    ctypes.byref(tmax)) == 0
    assert stride.value == n and tmax.value == 120 and tbp.value


def test_playout_totals_match_per_game_results():
    """othello_playout_args.totals = what play_a_game reports per game, summed over the launch"""
    w = torch.from_numpy(np.tile(np.arange(10, dtype=np.float32) - 3, (4, 1)).copy()).to(DEV)
    for kw in (dict(n_games=5000 + 13), dict(n_games=3000 + 5, policy=ops.POLICY_GREEDY, random_plies=6, weights=w)):
        tot = torch.zeros(4, dtype=torch.int64, device=DEV)
        po = ops.playout(seed=9, gid0=77, device=DEV, trajectory=False, totals=tot, **kw)
        c = po.final_counts().cpu().numpy().astype(np.int64)
        want = [int(po.nplies.sum()), int((c[:, 0] - c[:, 1]).sum()), int((c[:, 0] > c[:, 1]).sum()),
                int((c[:, 1] > c[:, 0]).sum())]
        assert tot.cpu().tolist() == want
        ops.playout(seed=9, gid0=77, device=DEV, trajectory=False, totals=tot, **kw)     # += semantics
        assert tot.cpu().tolist() == [2 * v for v in want]


def test_playout_host_async_two_batches_in_flight(ctx):
    """issue batch i+1 before waiting for batch i; every batch must equal the single-launch path and its
    totals the sums of its per-game results; outputs may be left out (NULL) individually"""
    L = _lib.lib()
    n = 200000 + 3
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    bufs = [(pin(n, torch.int32), pin(n, torch.int64), pin(n, torch.int64), pin(4, torch.int64), pin(n, torch.int16))
            for _ in range(2)]
    T = lambda t: ctypes.c_void_p(t.data_ptr())
    b0 = pin(n, torch.int64).fill_(ops.signed64(ops.START_BLACK))
    w0 = pin(n, torch.int64).fill_(ops.signed64(ops.START_WHITE))
    tickets = [0, 0]
    steps = 5
    for i in range(steps + 1):
        if i < steps:
            o = bufs[i % 2]
            tk = ctypes.c_int64()
            up = (T(b0), T(w0)) if i % 2 else (None, None)        # with and without uploaded start positions
            assert L.othello_playout_host_async(ctx, 21, 1000 * i, n, up[0], up[1], None, 0, 0, 0, 0, None, -1, None, 120,
                                                None, None, None, T(o[0]), T(o[1]), T(o[2]), T(o[4]), T(o[3]), ctypes.byref(tk)) == 0
            assert tk.value > 0
            tickets[i % 2] = tk.value
        if i > 0:
            j = i - 1
            assert L.othello_ctx_wait(ctx, tickets[j % 2]) == 0
            o = bufs[j % 2]
            po = ops.playout(n, seed=21, gid0=1000 * j, device=DEV, trajectory=False)
            assert torch.equal(o[0], po.nplies.cpu()) and torch.equal(o[1], po.final_black.cpu())
            assert torch.equal(o[2], po.final_white.cpu())
            c = po.final_counts().cpu().numpy().astype(np.int64)
            assert o[3].tolist() == [int(po.nplies.sum()), int((c[:, 0] - c[:, 1]).sum()),
                                     int((c[:, 0] > c[:, 1]).sum()), int((c[:, 1] > c[:, 0]).sum())]
            # the two-byte summary: plies in the low byte, n_black - n_white (int8) in the high byte
            sm = o[4].numpy().view(np.uint16)
            assert np.array_equal(sm & 0xff, po.nplies.cpu().numpy()) and \
                np.array_equal((sm >> 8).astype(np.uint8).view(np.int8), (c[:, 0] - c[:, 1]).astype(np.int8))
    assert L.othello_ctx_wait(ctx, tickets[0]) == 0 and L.othello_ctx_wait(ctx, 0) == 0      # waiting twice is harmless
    assert L.othello_ctx_wait(ctx, 10 ** 9) == -1                                            # a ticket never issued
    # totals only
    tot = pin(4, torch.int64)
    tk = ctypes.c_int64()
    assert L.othello_playout_host_async(ctx, 21, 0, n, None, None, None, 0, 0, 0, 0, None, -1, None, 120, None, None, None,
                                        None, None, None, None, T(tot), ctypes.byref(tk)) == 0
    assert L.othello_ctx_wait(ctx, tk.value) == 0
    assert int(tot[0]) == int(ops.playout(n, seed=21, gid0=0, device=DEV, trajectory=False).nplies.sum())
    assert L.othello_ctx_set_option(ctx, 1, 4) == 0 and L.othello_ctx_set_option(ctx, 99, 1) == -1
    # the synchronous small-batch calls still work while nothing is pending, and after a playout
    own = np.array([ops.START_BLACK], np.uint64); opp = np.array([ops.START_WHITE], np.uint64); out = np.zeros(1, np.uint64)
    assert L.othello_legal_host(ctx, P(own), P(opp), P(out), 1) == 0 and int(out[0]) == 0x0000102004080000


def test_three_batches_issued_before_any_wait(ctx):
    """a third batch reuses the first one's slot: the call completes the first, and its old ticket still waits fine"""
    L = _lib.lib()
    n = 70000
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    T = lambda t: ctypes.c_void_p(t.data_ptr())
    outs = [pin(n, torch.int32) for _ in range(3)]
    tks = []
    for i in range(3):
        tk = ctypes.c_int64()
        assert L.othello_playout_host_async(ctx, 5, 10 * i, n, None, None, None, 0, 0, 0, 0, None, -1, None, 120, None, None,
                                            None, T(outs[i]), None, None, None, None, ctypes.byref(tk)) == 0
        tks.append(tk.value)
    for i in (0, 2, 1):
        assert L.othello_ctx_wait(ctx, tks[i]) == 0
        assert torch.equal(outs[i], ops.playout(n, seed=5, gid0=10 * i, device=DEV, trajectory=False).nplies.cpu())
