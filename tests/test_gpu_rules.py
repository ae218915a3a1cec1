"""Parity of the rules / feature / evaluation kernels, through the C ABI, on a B200.

Bit-exact against (1) the golden vectors produced by the reference's own board.py and (2) the
pinned oracle on ~100k reachable positions, plus the edge cases (empty / ragged batches, full and
empty boards, occupied squares, unparsable hands, passes)."""
import numpy as np
import pytest
import torch

from subproc_b200 import ops
from gpu_util import DEV, dev_bits, dev_u8, dev_i32, host_bits, h, sample_positions

pytestmark = pytest.mark.gpu


def test_golden_games_every_ply(golden_games):
    for g in golden_games:
        pos, plies = g['positions'], g['plies']
        b = np.array([h(p['b']) for p in pos], dtype=np.uint64)
        w = np.array([h(p['w']) for p in pos], dtype=np.uint64)
        db, dw = dev_bits(b), dev_bits(w)
        assert host_bits(ops.legal(db, dw)).tolist() == [h(p['legal_b']) for p in pos]
        assert host_bits(ops.legal(dw, db)).tolist() == [h(p['legal_w']) for p in pos]
        assert ops.counts(db, dw).cpu().tolist() == [[p['nb'], p['nw'], p['ne']] for p in pos]
        if 'feat_O' in pos[0]:
            n = len(pos)
            assert ops.features(db, dw, dev_u8(np.full(n, 1))).cpu().tolist() == [p['feat_O'] for p in pos]
            assert ops.features(db, dw, dev_u8(np.full(n, 2))).cpu().tolist() == [p['feat_X'] for p in pos]
        n = len(plies)
        sb, sw = dev_bits(b[:n]), dev_bits(w[:n])
        turn = dev_u8([p['turn'] for p in pos[:n]])
        nturn = dev_i32([p['nturn'] for p in pos[:n]])
        fl, ret, flags = ops.step(sb, sw, turn, nturn, dev_u8([p['move'] for p in plies]))
        assert ret.cpu().tolist() == [p['ret'] for p in plies]
        assert host_bits(fl).tolist() == [h(p['flips']) for p in plies]
        assert host_bits(sb).tolist() == b[1:].tolist() and host_bits(sw).tolist() == w[1:].tolist()
        assert turn.cpu().tolist() == [p['turn'] for p in pos[1:]]
        assert nturn.cpu().tolist() == [p['nturn'] for p in pos[1:]]
        want_flags = []
        for p in pos[1:]:
            mover_legal = h(p['legal_b']) if p['turn'] == 1 else h(p['legal_w'])
            want_flags.append(ops.F_GAME_OVER if p['over'] else (ops.F_MUST_PASS if mover_legal == 0 else 0))
        assert flags.cpu().tolist() == want_flags


def test_golden_probe_put_both_colours(golden_probe):
    sq = dev_u8(np.arange(64))
    for rec in golden_probe:
        b0, w0 = h(rec['b']), h(rec['w'])
        db, dw = dev_bits(np.full(64, b0, np.uint64)), dev_bits(np.full(64, w0, np.uint64))
        for side in rec['put']:
            own, opp = (db, dw) if side['piece'] == 1 else (dw, db)
            fl = host_bits(ops.flips(own, opp, sq))
            assert fl.tolist() == [h(s) for s in side['flips']]
            assert [bin(int(v)).count('1') for v in fl] == side['ret']
        m = h(rec['mask'])
        got = ops.mask_count(db[:2].contiguous(), dw[:2].contiguous(), dev_u8([1, 2]), dev_bits([m, m])).cpu().tolist()
        assert got == rec['mask_count']
        sb, sw = db.clone(), dw.clone()
        turn, nturn = dev_u8(np.full(64, rec['turn'])), dev_i32(np.zeros(64))
        _, ret, _ = ops.step(sb, sw, turn, nturn, sq)
        assert ret.cpu().tolist() == rec['put_s']
        # illegal hands leave the position, turn and nturn untouched (board.py:199-201)
        bad = np.array(rec['put_s']) < 0
        assert (host_bits(sb)[bad] == b0).all() and (host_bits(sw)[bad] == w0).all()
        assert (turn.cpu().numpy()[bad] == rec['turn']).all() and (nturn.cpu().numpy()[bad] == 0).all()
        assert (nturn.cpu().numpy()[~bad] == 1).all()


def test_100k_reachable_positions_against_oracle(oracle):
    b, w = sample_positions(oracle, 1700, seed=21)
    n = b.size
    assert n > 100000
    db, dw = dev_bits(b), dev_bits(w)
    rng = np.random.RandomState(5)
    for piece in (1, 2):
        own, opp = (db, dw) if piece == 1 else (dw, db)
        assert np.array_equal(host_bits(ops.legal(own, opp)), oracle.puttables(b, w, piece))
        sq = rng.randint(0, 64, size=n).astype(np.uint8)
        _, _, want_fl, want_ret = oracle.put(b, w, piece, sq)
        got = host_bits(ops.flips(own, opp, dev_u8(sq)))
        assert np.array_equal(got, want_fl)
    assert np.array_equal(ops.counts(db, dw).cpu().numpy(), oracle.counts(b, w))
    side = rng.randint(1, 3, size=n).astype(np.uint8)
    assert np.array_equal(ops.features(db, dw, dev_u8(side)).cpu().numpy(), oracle.features(b, w, side))

    # put_s with a mix of legal moves, illegal squares, passes and unparsable hands
    turn = rng.randint(1, 3, size=n).astype(np.uint8)
    nturn = rng.randint(0, 100, size=n).astype(np.int32)
    legal_turn = np.where(turn == 1, oracle.puttables(b, w, 1), oracle.puttables(b, w, 2))
    move = rng.randint(0, 64, size=n).astype(np.uint8)
    pick_legal = rng.rand(n) < 0.6
    for i in np.nonzero(pick_legal & (legal_turn != 0))[0]:
        bits = [s for s in range(64) if (int(legal_turn[i]) >> s) & 1]
        move[i] = bits[rng.randint(len(bits))]
    move[rng.rand(n) < 0.05] = 64
    move[rng.rand(n) < 0.02] = rng.randint(65, 256)
    wb, ww, wt, wnt, wfl, wret = oracle.step(b, w, turn, nturn, move)
    sb, sw, st, snt = dev_bits(b), dev_bits(w), dev_u8(turn), dev_i32(nturn)
    fl, ret, flags = ops.step(sb, sw, st, snt, dev_u8(move))
    assert np.array_equal(ret.cpu().numpy(), wret)
    assert np.array_equal(host_bits(fl), wfl)
    assert np.array_equal(host_bits(sb), wb) and np.array_equal(host_bits(sw), ww)
    assert np.array_equal(st.cpu().numpy(), wt) and np.array_equal(snt.cpu().numpy(), wnt)
    over = oracle.game_over(wb, ww)
    mover_legal = np.where(wt == 1, oracle.puttables(wb, ww, 1), oracle.puttables(wb, ww, 2))
    want_flags = np.where(over == 1, ops.F_GAME_OVER, np.where(mover_legal == 0, ops.F_MUST_PASS, 0))
    assert np.array_equal(flags.cpu().numpy(), want_flags.astype(np.uint8))
    assert (wret > 0).sum() > 30000 and (wret == 0).sum() > 1000 and (wret < 0).sum() > 10000


def test_eval_tolerance_float_weights_and_exact_integer_weights(oracle):
    """north_star: evaluation within 1e-5 relative for floating-point weights; integer weights
    (the stored form, progress_position_moves_learn.py:200) are exact."""
    b, w = sample_positions(oracle, 300, seed=22)
    n = b.size
    db, dw = dev_bits(b), dev_bits(w)
    rng = np.random.RandomState(9)
    side = rng.randint(1, 3, size=n).astype(np.uint8)
    wf = np.concatenate([rng.uniform(-3, 3, size=(4, 9)), rng.uniform(-10, 10, size=(4, 1))], axis=1)
    got = ops.evaluate(db, dw, dev_u8(side), torch.from_numpy(wf.astype(np.float32)).to(DEV)).cpu().numpy()
    want = oracle.evaluate(b, w, side, wf.astype(np.float32).astype(np.float64))
    feats = oracle.features(b, w, side)[:, 1:].astype(np.float64)
    rows = np.minimum(np.maximum((oracle.features(b, w, side)[:, 0] + 15) // 16 - 1, 0), 3)
    scale = (np.abs(wf[rows, :9]) * feats).sum(axis=1) + np.abs(wf[rows, 9]) + 1e-30
    assert np.max(np.abs(got - want) / scale) <= 1e-5          # tolerance stated by BASELINE.json north_star
    wi = oracle.DEFAULT_WEIGHTS
    got = ops.evaluate(db, dw, dev_u8(side), torch.from_numpy(wi.astype(np.float32)).to(DEV)).cpu().numpy()
    assert np.array_equal(got.astype(np.float64), oracle.evaluate(b, w, side, wi))


def test_edge_cases_empty_ragged_full_and_empty_boards(oracle):
    e = torch.empty(0, dtype=torch.int64, device=DEV)
    assert ops.legal(e, e).numel() == 0
    assert ops.counts(e, e).shape == (0, 3)
    # ragged sizes around the block size
    b, w = sample_positions(oracle, 8, seed=23)
    for n in (1, 31, 255, 256, 257):
        assert np.array_equal(host_bits(ops.legal(dev_bits(b[:n]), dev_bits(w[:n]))), oracle.puttables(b[:n], w[:n], 1))
    full = 0xFFFFFFFFFFFFFFFF
    cases_b = np.array([0, full, 0, full & 0x5555555555555555, 1, 0x8000000000000000], dtype=np.uint64)
    cases_w = np.array([0, 0, full, full & 0xAAAAAAAAAAAAAAAA, 0x8000000000000000, 1], dtype=np.uint64)
    db, dw = dev_bits(cases_b), dev_bits(cases_w)
    assert np.array_equal(host_bits(ops.legal(db, dw)), oracle.puttables(cases_b, cases_w, 1))
    assert np.array_equal(host_bits(ops.legal(dw, db)), oracle.puttables(cases_b, cases_w, 2))
    assert np.array_equal(ops.counts(db, dw).cpu().numpy(), oracle.counts(cases_b, cases_w))
    turn, nturn = dev_u8(np.ones(6)), dev_i32(np.zeros(6))
    _, ret, flags = ops.step(db, dw, turn, nturn, dev_u8([0, 0, 64, 5, 64, 200]))
    assert ret.cpu().tolist() == [-1, -1, 0, -1, 0, -1]
    assert flags.cpu().tolist() == [2, 2, 2, 2, 2, 2]
    with pytest.raises(ValueError):
        ops.legal(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64))     # CPU tensors: no fallback


def test_edge_wraparound_lines(oracle):
    """runs that touch the a/h files and ranks 1/8 must not wrap (is_within_board, board.py:131-137)."""
    rng = np.random.RandomState(31)
    n = 20000
    occ = rng.randint(0, 2 ** 62, size=n).astype(np.uint64) | (rng.randint(0, 4, size=n).astype(np.uint64) << np.uint64(62))
    col = rng.randint(0, 2 ** 62, size=n).astype(np.uint64) | (rng.randint(0, 4, size=n).astype(np.uint64) << np.uint64(62))
    b, w = occ & col, occ & ~col                                  # dense random (unreachable) positions
    assert np.array_equal(host_bits(ops.legal(dev_bits(b), dev_bits(w))), oracle.puttables(b, w, 1))
    sq = rng.randint(0, 64, size=n).astype(np.uint8)
    _, _, want_fl, _ = oracle.put(b, w, 2, sq)
    assert np.array_equal(host_bits(ops.flips(dev_bits(w), dev_bits(b), dev_u8(sq))), want_fl)


def test_constructed_lines_every_direction_square_and_run_length(oracle):
    """both formulations on the GPU (plain Kogge-Stone in rules.cu, tuned in the game kernels) on runs
    of 0..7 discs from every square in every direction, closed by an own disc / open / into the edge"""
    import line_cases
    own, opp, sq = line_cases.build()
    d_own, d_opp = dev_bits(own), dev_bits(opp)
    assert np.array_equal(host_bits(ops.legal(d_own, d_opp)), oracle.puttables(own, opp, 1))
    _, _, want, ret = oracle.put(own, opp, 1, sq)
    assert np.array_equal(host_bits(ops.flips(d_own, d_opp, dev_u8(sq))), want)
    # the tuned primitives, through a one-ply greedy search from each position: the chosen move's
    # successor must equal the oracle's greedy successor
    n = own.size
    w = torch.from_numpy(oracle.DEFAULT_WEIGHTS.astype(np.float32)).to(DEV)
    has_move = oracle.puttables(own, opp, 1) != 0
    idx = np.nonzero(has_move)[0][:4000]
    po = ops.playout(idx.size, seed=1, gid0=0, device=DEV, black0=dev_bits(own[idx]), white0=dev_bits(opp[idx]),
                     policy=ops.POLICY_GREEDY, weights=w, t_max=2)
    ref = oracle.playout(1, 0, idx.size, black0=own[idx], white0=opp[idx], policy=1, t_max=2)
    assert np.array_equal(host_bits(po.black[1]), ref['black'][1]) and np.array_equal(host_bits(po.white[1]), ref['white'][1])
    assert np.array_equal(po.move[0].cpu().numpy(), ref['move'][0])
