"""The reference's GameRunner (game_runner.py:105-201) over the lock-step playout kernel.

The reference plays ONE game between two engine subprocesses; here the "engines" are policies
evaluated inside the playout kernels, and ``play_games(n)`` plays n games in one launch.  The
constructor keeps the reference's argument order (game_runner.py:107): where the reference takes
two shell commands, this takes two engine specs -- Black's and White's may differ (random vs greedy,
or two greedy engines with different parameter sets), like proc_black / proc_white.

The recorder protocol is the reference's: ``recorder.add(board)`` for the initial position and
after every ply (game_runner.py:170,159), ``add_meta`` + ``store`` at the end (:186-192).  Boards
handed to recorders are ``subproc_b200.board.Board`` objects rebuilt from the trajectory.
"""
import torch

from . import ops
from . import board as board_mod

N_RAND_HAND_UNTIL = 10                               # game_runner.py:6 (applied inside the kernel)


class Engine(object):
    """What sits behind Player.go (game_runner.py:19-33): 'random' or 'greedy' on a weight table."""

    def __init__(self, policy='random', weights=None, random_plies=0, name=None):
        if policy not in ('random', 'greedy'):
            raise ValueError("engine policy must be 'random' or 'greedy'")
        self.policy = policy
        self.weights = weights                       # 4 x 9 (default_value() layout) or 4 x 10
        self.random_plies = random_plies
        self.name = name or ('b200-' + policy)


class GameRunner(object):
    def __init__(self, proc_black, proc_white, game_recorder=None, debug=False, n_rand_hands_for_black=0,
                 n_rand_hands_for_white=0, device=None, seed=0):
        if not isinstance(proc_black, Engine):
            proc_black = Engine(proc_black)
        if not isinstance(proc_white, Engine):
            proc_white = Engine(proc_white)
        self.proc_black, self.proc_white = proc_black, proc_white
        self.recorder = game_recorder
        self.debug = debug
        self.n_rand_black = int(n_rand_hands_for_black)
        self.n_rand_white = int(n_rand_hands_for_white)
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.seed = seed
        self.next_gid = 0
        greedy = [e for e in (proc_black, proc_white) if e.policy == 'greedy']
        if len(greedy) == 2 and greedy[0].random_plies != greedy[1].random_plies:
            raise ValueError("the two greedy engines must share random_plies (one opening length per game)")
        self.random_plies = greedy[0].random_plies if greedy else 0
        self._weights = [self._table(e) for e in (proc_black, proc_white)]

    def _table(self, engine):
        if engine.policy != 'greedy':
            return None
        import numpy as np
        w = np.asarray(engine.weights, dtype=np.float64).reshape(4, -1)
        return ops.weights_tensor(w[:, :9], self.device, w[:, 9] if w.shape[1] > 9 else None)

    def play_games(self, n_games, trajectory=True, t_max=ops.T_MAX_DEFAULT, out=None):
        """n_games games in one launch; returns an ops.Playout (trajectories stay in HBM)."""
        pol = [ops.POLICY_GREEDY if e.policy == 'greedy' else ops.POLICY_RANDOM for e in (self.proc_black, self.proc_white)]
        po = ops.playout(n_games, seed=self.seed, gid0=self.next_gid, device=self.device, policy=pol[0],
                         policy_white=pol[1], random_plies=self.random_plies, n_rand_black=self.n_rand_black,
                         n_rand_white=self.n_rand_white, weights=self._weights[0], weights_white=self._weights[1],
                         t_max=t_max, trajectory=trajectory, out=out)
        self.next_gid += int(n_games)
        return po

    def winners(self, po):
        """int8[n]: +1 Black won, -1 White won, 0 draw (game_runner.py:194-199)."""
        c = po.final_counts()
        return torch.sign(c[:, 0] - c[:, 1]).to(torch.int8)

    def play_a_game(self):
        """One game with the reference's return value ((colour, engine name)) and recorder calls."""
        po = self.play_games(1)
        n = int(po.nplies.cpu()[0])
        if self.recorder is not None:
            blacks = ops.bits_numpy(po.black[:n + 1, 0])
            whites = ops.bits_numpy(po.white[:n + 1, 0])
            turn = board_mod.Black
            for t in range(n + 1):
                b = board_mod.Board(device=self.device)
                b._black, b._white, b.turn, b.nturn = int(blacks[t]), int(whites[t]), turn, t
                self.recorder.add(b)
                turn = b.hostile(turn)
            self.recorder.add_meta({'proc_a': self.proc_black.name, 'proc_b': self.proc_white.name,
                                    'hamletparam': 'No Hamlet'})
            self.recorder.store()
        c = po.final_counts().cpu()
        nb, nw = int(c[0, 0]), int(c[0, 1])
        if nb > nw:
            return ("Black", self.proc_black.name)
        if nb < nw:
            return ("White", self.proc_white.name)
        return ("None", '')
