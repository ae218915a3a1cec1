"""BatchedOthello (Board semantics for B games in HBM) and the batched GameRunner on a B200."""
import numpy as np
import pytest
import torch

from subproc_b200 import ops
from subproc_b200.batched import BatchedOthello
from subproc_b200.game_runner import GameRunner, Engine
from gpu_util import DEV, host_bits

pytestmark = pytest.mark.gpu


def test_batched_games_step_by_step_against_oracle(oracle):
    """play 512 games ply by ply through put_s with moves drawn on the host from the legal masks"""
    n = 512
    env = BatchedOthello(n, device=DEV)
    rng = np.random.RandomState(4)
    b = np.full(n, oracle.START_BLACK, np.uint64)
    w = np.full(n, oracle.START_WHITE, np.uint64)
    turn = np.ones(n, np.uint8)
    nturn = np.zeros(n, np.int32)
    assert not env.is_game_over().any()
    for ply in range(130):
        legal = host_bits(env.puttables_for_turn())
        want_legal = np.where(turn == 1, oracle.puttables(b, w, 1), oracle.puttables(b, w, 2))
        assert np.array_equal(legal, want_legal)
        move = np.full(n, 64, np.uint8)
        for i in np.nonzero(legal)[0]:
            bits = [s for s in range(64) if (int(legal[i]) >> s) & 1]
            move[i] = bits[rng.randint(len(bits))]
        if ply % 7 == 3:
            move[::5] = rng.randint(0, 64, size=move[::5].size)          # sprinkle (mostly) illegal hands
        ret = env.put_s(torch.from_numpy(move).to(DEV)).cpu().numpy()
        b, w, turn, nturn, fl, want_ret = oracle.step(b, w, turn, nturn, move)
        assert np.array_equal(ret, want_ret)
        assert np.array_equal(host_bits(env.black), b) and np.array_equal(host_bits(env.white), w)
        assert np.array_equal(env.turn.cpu().numpy(), turn) and np.array_equal(env.nturn.cpu().numpy(), nturn)
        assert np.array_equal(host_bits(env.last_flips()), fl)
        over = oracle.game_over(b, w).astype(bool)
        assert np.array_equal(env.is_game_over().cpu().numpy(), over)
        assert np.array_equal((env.flags.cpu().numpy() & ops.F_GAME_OVER) != 0, over)
        if over.all():
            break
    assert over.all()
    c = env.counts().cpu().numpy()
    assert np.array_equal(c, oracle.counts(b, w))
    assert np.array_equal(env.n_puttable_for(1).cpu().numpy(), np.zeros(n, np.int32))
    assert np.array_equal(env.features(2).cpu().numpy(), oracle.features(b, w, 2))
    assert env.serialize_board(0) == ''.join('O' if (int(b[0]) >> s) & 1 else 'X' if (int(w[0]) >> s) & 1 else '-' for s in range(64))


def test_game_runner_batches_and_winners(oracle):
    gr = GameRunner(Engine('greedy', oracle.DEFAULT_WEIGHTS, random_plies=6), Engine('greedy', oracle.DEFAULT_WEIGHTS, random_plies=6),
                    None, False, 2, 3, device=DEV, seed=21)
    po1 = gr.play_games(300)
    po2 = gr.play_games(200)                                     # game ids continue: 300..499
    ref = oracle.playout(21, 0, 500, policy=1, random_plies=6, n_rand_black=2, n_rand_white=3)
    got = np.concatenate([po1.nplies.cpu().numpy(), po2.nplies.cpu().numpy()])
    assert np.array_equal(got, ref['nplies'])
    fb = np.concatenate([host_bits(po1.final_black), host_bits(po2.final_black)])
    assert np.array_equal(fb, ref['final_black'])
    win = gr.winners(po1).cpu().numpy()
    cnt = oracle.counts(ref['final_black'][:300], ref['final_white'][:300])
    assert np.array_equal(win, np.sign(cnt[:, 0] - cnt[:, 1]).astype(np.int8))
    # different engines per colour, like proc_black / proc_white (game_runner.py:107-123)
    other = np.array([[3, 80, 40, -20, 5, 5, 1, 1, 2, 0], [10, 60, 30, -10, 4, 6, 2, 2, 1, 0],
                      [20, 50, 20, -5, 3, 7, 3, 3, 3, 0], [64, 10, 10, 10, 10, 10, 10, 10, 10, 0]], dtype=np.float64)
    for black, white, kw in (
            ('random', Engine('greedy', oracle.DEFAULT_WEIGHTS), dict(policy=0, policy_white=1)),
            (Engine('greedy', other, random_plies=3), 'random', dict(policy=1, policy_white=0, weights=other, random_plies=3)),
            (Engine('greedy', oracle.DEFAULT_WEIGHTS, random_plies=2), Engine('greedy', other, random_plies=2),
             dict(policy=1, policy_white=1, weights_white=other, random_plies=2))):
        gr = GameRunner(black, white, None, False, 1, 2, device=DEV, seed=22)
        po = gr.play_games(400)
        ref = oracle.playout(22, 0, 400, n_rand_black=1, n_rand_white=2, **kw)
        assert np.array_equal(po.nplies.cpu().numpy(), ref['nplies'])
        assert np.array_equal(host_bits(po.final_black), ref['final_black'])
        assert np.array_equal(host_bits(po.final_white), ref['final_white'])
    with pytest.raises(ValueError):
        GameRunner(Engine('greedy', other, random_plies=1), Engine('greedy', other, random_plies=2), None, False, 0, 0, device=DEV)
