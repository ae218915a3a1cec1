"""Parity of the lock-step playout kernel (GameRunner.play_a_game) and perft on a B200."""
import numpy as np
import pytest
import torch

from subproc_b200 import ops
from gpu_util import DEV, dev_bits, dev_u8, dev_i32, host_bits, h, sample_positions

pytestmark = pytest.mark.gpu


def check_against_oracle(po, ref, n):
    nplies = po.nplies.cpu().numpy()
    assert np.array_equal(nplies, ref['nplies'])
    assert np.array_equal(host_bits(po.final_black), ref['final_black'])
    assert np.array_equal(host_bits(po.final_white), ref['final_white'])
    tb, tw, tm = host_bits(po.black), host_bits(po.white), po.move.cpu().numpy()
    t_idx = np.arange(po.t_max + 1)[:, None]
    valid_pos = t_idx <= np.minimum(nplies, po.t_max)[None, :]
    valid_mv = t_idx[:-1] < np.minimum(nplies, po.t_max)[None, :]
    assert np.array_equal(tb[valid_pos], ref['black'][valid_pos])
    assert np.array_equal(tw[valid_pos], ref['white'][valid_pos])
    assert np.array_equal(tm[valid_mv], ref['move'][valid_mv])


def test_golden_games_replayed_by_the_kernel(golden_games):
    """the games the reference's board.py played in oracle/make_golden.py, move for move"""
    for g in golden_games:
        w = ops.weights_tensor(np.array(
            [[100, 99, -1, -1, -1, -1, 3, 8, 20], [75, 99, 2, -5, 7, 6, 4, 5, 5],
             [25, 99, 2, -5, -7, -6, 4, 5, 5], [1, 100, 50, 30, 30, 30, 30, 30, 30]]), DEV)
        ww = ops.weights_tensor(np.array(g['rows_white']), DEV) if 'rows_white' in g else None
        po = ops.playout(1, seed=g['seed'], gid0=g['gid'], device=DEV, policy=g['policy'],
                         random_plies=g['random_plies'], n_rand_black=g['n_rand_black'],
                         n_rand_white=g['n_rand_white'], weights=w, policy_white=g.get('policy_white'),
                         weights_white=ww)
        n = len(g['plies'])
        assert int(po.nplies.cpu()[0]) == n
        assert po.move[:n, 0].cpu().tolist() == [p['move'] for p in g['plies']]
        assert host_bits(po.black[:n + 1, 0].contiguous()).tolist() == [h(p['b']) for p in g['positions']]
        assert host_bits(po.white[:n + 1, 0].contiguous()).tolist() == [h(p['w']) for p in g['positions']]


def test_random_playouts_bit_exact(oracle):
    n = 4096 + 37                                      # ragged: not a multiple of the block size
    po = ops.playout(n, seed=1, gid0=0, device=DEV)
    check_against_oracle(po, oracle.playout(1, 0, n), n)
    assert 55 < po.nplies.float().mean().item() < 65


def test_substitution_rule_and_greedy_bit_exact(oracle):
    w = torch.from_numpy(oracle.DEFAULT_WEIGHTS.astype(np.float32)).to(DEV)
    n = 700
    po = ops.playout(n, seed=5, gid0=100, device=DEV, n_rand_black=3, n_rand_white=10)
    check_against_oracle(po, oracle.playout(5, 100, n, n_rand_black=3, n_rand_white=10), n)
    po = ops.playout(n, seed=6, gid0=0, device=DEV, policy=ops.POLICY_GREEDY, random_plies=10, weights=w)
    check_against_oracle(po, oracle.playout(6, 0, n, policy=1, random_plies=10), n)
    po = ops.playout(n, seed=7, gid0=0, device=DEV, policy=ops.POLICY_GREEDY, n_rand_black=10, n_rand_white=2, weights=w)
    check_against_oracle(po, oracle.playout(7, 0, n, policy=1, n_rand_black=10, n_rand_white=2), n)


def test_custom_start_positions_and_turns(oracle):
    b, w = sample_positions(oracle, 40, seed=41, stride=3)
    n = b.size
    turn = (np.arange(n) % 2 + 1).astype(np.uint8)
    po = ops.playout(n, seed=9, gid0=7, device=DEV, black0=dev_bits(b), white0=dev_bits(w), turn0=dev_u8(turn))
    check_against_oracle(po, oracle.playout(9, 7, n, black0=b, white0=w, turn0=turn), n)


def test_sharding_invariance_and_no_trajectory_mode(oracle):
    """game g depends only on (seed, gid): any split over launches / GPUs gives the same games"""
    n = 3000
    whole = ops.playout(n, seed=3, gid0=0, device=DEV)
    a = ops.playout(1000, seed=3, gid0=0, device=DEV)
    b = ops.playout(2000, seed=3, gid0=1000, device=DEV, trajectory=False)
    assert torch.equal(whole.nplies, torch.cat([a.nplies, b.nplies]))
    assert torch.equal(whole.final_black, torch.cat([a.final_black, b.final_black]))
    assert torch.equal(whole.final_white, torch.cat([a.final_white, b.final_white]))
    live = torch.arange(whole.t_max, device=DEV)[:, None] < a.nplies[None, :]      # rows past a game's end are unspecified
    assert torch.equal(whole.move[:, :1000][live], a.move[live])
    assert torch.equal(whole.black[:-1, :1000][live], a.black[:-1][live])
    assert b.black is None


def test_truncated_trajectory_capacity(oracle):
    n = 500
    po = ops.playout(n, seed=2, gid0=0, device=DEV, t_max=20)
    ref = oracle.playout(2, 0, n, t_max=20)
    check_against_oracle(po, ref, n)
    assert (po.nplies > 20).all()


def test_zero_capacity_and_single_game(oracle):
    po = ops.playout(5, seed=2, gid0=0, device=DEV, t_max=0)              # records the start position only
    assert np.array_equal(po.nplies.cpu().numpy(), oracle.playout(2, 0, 5)['nplies'])
    assert host_bits(po.black[0]).tolist() == [oracle.START_BLACK] * 5
    one = ops.playout(1, seed=2, gid0=3, device=DEV)
    assert int(one.nplies[0]) == int(po.nplies[3])
    b0, w0, t0 = [0, 1, 2 ** 64 - 1, 1], [0, 2, 0, 4], [1, 2, 1, 1]           # empty, pass-then-move, full, dead
    tiny = ops.playout(4, seed=1, gid0=0, device=DEV, black0=dev_bits(b0), white0=dev_bits(w0), turn0=dev_u8(t0))
    ref = oracle.playout(1, 0, 4, black0=np.array(b0, np.uint64), white0=np.array(w0, np.uint64), turn0=np.array(t0, np.uint8))
    assert tiny.nplies.cpu().tolist() == ref['nplies'].tolist() == [0, 2, 0, 0]
    assert np.array_equal(host_bits(tiny.final_black), ref['final_black'])
    assert np.array_equal(host_bits(tiny.final_white), ref['final_white'])


def test_million_games_round_trip_properties():
    """BASELINE config 3 at full size: every recorded ply, replayed through the step kernel, must
    reproduce the next recorded position; finals are terminal; disc counts are consistent."""
    n = 1 << 20
    po = ops.playout(n, seed=1, gid0=0, device=DEV)
    nplies = po.nplies
    assert int(nplies.min()) >= 9 and int(nplies.max()) <= 120
    total = po.total_positions()
    assert 60.0 < total / n < 61.0
    fin = ops.legal(po.final_black, po.final_white) | ops.legal(po.final_white, po.final_black)
    assert int((fin != 0).sum()) == 0                                  # finals are game-over positions
    c = po.final_counts()
    assert int((c.sum(dim=1) != 64).sum()) == 0 and int(c[:, 2].max()) <= 60
    turn = torch.ones(n, dtype=torch.uint8, device=DEV)
    nturn = torch.zeros(n, dtype=torch.int32, device=DEV)
    for t in (0, 1, 7, 30, 52, 58, 59, 60, 61):
        live = nplies > t
        b, w = po.black[t].clone(), po.white[t].clone()
        turn.fill_(1 if t % 2 == 0 else 2)
        nturn.fill_(t)
        _, ret, _ = ops.step(b, w, turn, nturn, po.move[t].contiguous())
        assert int((ret[live] < 0).sum()) == 0
        assert torch.equal(b[live], po.black[t + 1][live]) and torch.equal(w[live], po.white[t + 1][live])
    last_b = po.black.gather(0, nplies.long()[None, :])[0]
    assert torch.equal(last_b, po.final_black)


def test_perft_known_answers(oracle):
    want = [1, 4, 12, 56, 244, 1396, 8200, 55092, 390216, 3005288, 24571284]
    for d, v in enumerate(want):
        assert ops.perft(d, device=DEV) == v, d
    # published Othello perft values beyond the survey's table (same counting convention)
    assert ops.perft(11, device=DEV) == 212258800
    assert ops.perft(12, device=DEV) == 1939886636


def test_perft_other_roots_against_oracle(oracle):
    b, w = sample_positions(oracle, 6, seed=77, stride=9)
    for i in range(0, b.size, 2):
        for turn in (1, 2):
            for d in (1, 2, 4):
                assert ops.perft(d, int(b[i]), int(w[i]), turn, device=DEV) == oracle.perft(d, int(b[i]), int(w[i]), turn)


def test_perft_parts_add_up_and_workspace_follows_depth(oracle):
    """the depth-first stage split nparts ways (frontier split + one u64 all-reduce): the shares,
    left on the device by a launch sequence the host never waits for, add up to the node count"""
    from subproc_b200 import _lib
    L = _lib.lib()
    sizes = [int(L.othello_perft_workspace_bytes(d)) for d in (1, 2, 3, 4, 6, 10, 30)]
    assert sizes == sorted(sizes) and sizes[0] < 8192 and sizes[-1] == sizes[-2]      # grows with depth, then capped
    for depth, want in ((3, 56), (7, 55092), (10, 24571284)):
        for nparts in (1, 3, 8):
            shares = [ops.perft_part(depth, p, nparts, device=DEV) for p in range(nparts)]     # all enqueued, no sync
            tot = torch.stack(shares).sum(0).cpu().tolist()
            assert tot == [want, 0], (depth, nparts, tot)
    assert ops.perft_distributed(9, device=DEV) == 3005288                 # single process: world of one
    # an endgame root: game-over leaves inside the breadth-first levels are counted once (by part 0)
    r = oracle.playout(5, 0, 4)
    g = 0
    t = int(r['nplies'][g]) - 6
    b, w = int(r['black'][t, g]), int(r['white'][t, g])
    turn = 1 if t % 2 == 0 else 2
    for d in (5, 8, 12):
        want = oracle.perft(d, b, w, turn)
        assert ops.perft(d, b, w, turn, device=DEV) == want
        assert sum(int(ops.perft_part(d, p, 4, b, w, turn, device=DEV)[0]) for p in range(4)) == want
    # too little scratch is reported, not overrun
    ws = torch.empty(256 + 16 + 4 * 64 * 8, dtype=torch.uint8, device=DEV)
    import ctypes
    res = ctypes.c_uint64()
    rc = L.othello_perft(ops.START_BLACK, ops.START_WHITE, 1, 9, ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                         ctypes.byref(res), None)
    assert rc == -2


def test_greedy_games_do_not_depend_on_games_per_warp(oracle):
    """the greedy kernel's small-batch modes (8 / 16 games per warp, the other lanes only evaluate children)
    play the same games as the 32-games-per-warp layout and as the oracle"""
    w = torch.from_numpy(oracle.DEFAULT_WEIGHTS.astype(np.float32)).to(DEV)
    n = 3000 + 7
    ref = oracle.playout(19, 40, 600, policy=1, random_plies=6, n_rand_black=2, n_rand_white=1)
    outs = []
    for gpw in (32, 16, 8, 0):
        tot = torch.zeros(4, dtype=torch.int64, device=DEV)
        sm = torch.zeros(n, dtype=torch.int16, device=DEV)
        po = ops.playout(n, seed=19, gid0=40, device=DEV, policy=ops.POLICY_GREEDY, random_plies=6, n_rand_black=2,
                         n_rand_white=1, weights=w, games_per_warp=gpw, totals=tot, summary=sm)
        c = po.final_counts().cpu().numpy().astype(np.int64)
        smv = sm.cpu().numpy().view(np.uint16)
        assert np.array_equal(smv & 0xff, po.nplies.cpu().numpy())
        assert np.array_equal((smv >> 8).astype(np.uint8).view(np.int8), (c[:, 0] - c[:, 1]).astype(np.int8))
        assert np.array_equal(po.nplies[:600].cpu().numpy(), ref['nplies'])
        assert np.array_equal(ops.bits_numpy(po.final_black[:600]), ref['final_black'])
        t = 30
        live = ref['nplies'] > t
        assert np.array_equal(ops.bits_numpy(po.black[t][:600])[live], ref['black'][t][live])
        assert np.array_equal(po.move[t][:600].cpu().numpy()[live], ref['move'][t][live])
        outs.append((po.nplies.clone(), po.final_black.clone(), po.final_white.clone(), tot))
    for o in outs[1:]:
        assert all(torch.equal(x, y) for x, y in zip(o, outs[0]))


def test_greedy_with_float_weights_how_often_fp32_and_fp64_disagree(oracle):
    """Greedy games are bit-exact for integer-valued weights (the stored form).  With arbitrary float weights the
    device evaluates in fp32 and the oracle in fp64: scores agree to 1e-5 relative, so a game can only differ where
    two successors are closer than that.  Count it: weights with three decimals (ties at the 1e-3 level are then
    real ties in both precisions, anything closer is rounding), 4000 games of ~50 greedy plies each."""
    rng = np.random.RandomState(11)
    wts = np.zeros((4, 10))
    wts[:, :9] = np.round(rng.uniform(-60, 100, size=(4, 9)), 3)
    wts[:, 9] = np.round(rng.uniform(-5, 5, size=4), 3)
    n = 4000
    ref = oracle.playout(29, 0, n, policy=1, random_plies=10, weights=wts)
    po = ops.playout(n, seed=29, gid0=0, device=DEV, policy=ops.POLICY_GREEDY, random_plies=10,
                     weights=torch.from_numpy(wts.astype(np.float32)).to(DEV))
    same = (po.nplies.cpu().numpy() == ref['nplies']) & (ops.bits_numpy(po.final_black) == ref['final_black']) & \
           (ops.bits_numpy(po.final_white) == ref['final_white'])
    # first ply at which a differing game leaves the oracle's move list
    mv = po.move.cpu().numpy()
    first = [int(np.argmax(mv[:ref['nplies'][g], g] != ref['move'][:ref['nplies'][g], g])) for g in np.flatnonzero(~same)]
    print("float weights: %d of %d games identical; differing games first diverge at plies %s" % (same.sum(), n, sorted(first)[:10]))
    assert same.mean() >= 0.99
    assert all(t >= 10 for t in first)                         # never inside the random opening
