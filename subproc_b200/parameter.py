"""Drop-in for the reference's ``parameter`` / ``parameter_progress_position_moves_learn`` modules.

``counts(a_book, side)`` and ``ProgressPositionMovesParameter`` keep the reference's names, argument
meaning and return formats (parameter.py:5-63, parameter_progress_position_moves_learn.py:5-49); the
arithmetic (mobility via legal-move generation, eight square-class counts, disc count) runs in the
feature kernel.  Batched forms (``counts_batch``) take many books / bitboards per launch.
"""
from abc import ABCMeta, abstractmethod

import numpy as np

from . import board


def board_from_a_book(a_book):                          # parameter.py:5-8
    a_board = board.Board()
    a_board.deserialize(a_book['book'], a_book['whosturn'], a_book['turn'])
    return a_board


class ParameterBase(object, metaclass=ABCMeta):         # parameter.py:11-36
    @abstractmethod
    def configure(self, conf):
        pass

    @abstractmethod
    def header(self):
        pass

    @abstractmethod
    def default_value(self):
        pass

    @abstractmethod
    def features_from_hash(self, hash_key):
        pass

    @abstractmethod
    def phase_from_hash(self, hash_key):
        pass

    @abstractmethod
    def hash_from_book(self, a_book, side):
        pass


class ParameterBasePlus(ParameterBase):                 # parameter.py:39-63
    pass


def bits_from_book_string(book_str):
    """64-char 'O'/'X'/'-' string (board.py:223-243: 'O' = Black, 'X' = White) -> (black, white)."""
    black = white = 0
    for i, ch in enumerate(book_str):
        if ch == 'O':
            black |= 1 << i
        elif ch == 'X':
            white |= 1 << i
    return black, white


def counts_batch(blacks, whites, sides, device=None):
    """counts() for many positions in one launch: int32 [n][10] on the host.

    blacks/whites: iterables of bit patterns; sides: iterable of 'O'/'X' or colours 1/2."""
    import torch
    from . import ops
    dev = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
    col = [board.Black if s in ('O', board.Black) else board.White if s in ('X', board.White) else board.Empty
           for s in sides]
    out = ops.features(ops.bits_tensor(list(blacks), dev), ops.bits_tensor(list(whites), dev),
                       torch.tensor(col, dtype=torch.uint8, device=dev))
    return out.cpu().numpy()


def counts(a_book, side):                               # parameter_progress_position_moves_learn.py:5-17
    black, white = bits_from_book_string(a_book['book'])
    return tuple(int(v) for v in counts_batch([black], [white], [side])[0])


class ProgressPositionMovesParameter(ParameterBasePlus):    # parameter_progress_position_moves_learn.py:20-49
    def __init__(self):
        pass

    def configure(self, conf):
        pass

    def header(self):
        return 2

    def default_value(self):
        return [
            [100, 99, -1, -1, -1, -1, 3, 8, 20],
            [75, 99, 2, -5, 7, 6, 4, 5, 5],
            [25, 99, 2, -5, -7, -6, 4, 5, 5],
            [1, 100, 50, 30, 30, 30, 30, 30, 30]
        ]

    def _state_from_hash(self, hash_key):
        # keys look like 'othelloparam:<learner>:param:state:<10 ints>' (parameter_store.py:70-83)
        return tuple(int(x) for x in hash_key.split(":")[4:])

    def features_from_hash(self, hash_key):
        return self._state_from_hash(hash_key)[1:]

    def phase_from_hash(self, hash_key):
        return self._state_from_hash(hash_key)[0]

    def hash_from_book(self, a_book, side):
        return ':'.join(str(x) for x in counts(a_book, side))

    def hashes_from_books(self, books, sides):
        """batched hash_from_book: one launch for all (book, side) pairs."""
        bw = [bits_from_book_string(b['book']) for b in books]
        f = counts_batch([x[0] for x in bw], [x[1] for x in bw], sides)
        return [':'.join(str(int(v)) for v in row) for row in f]

    def weights_table(self, parameters=None):
        """[4][10] float32 table (9 weights + zero intercept per phase row) for the eval kernels, from
        default_value() or from the (header, w0..w35) tuple read_parameters returns."""
        if parameters is None:
            rows = np.asarray(self.default_value(), dtype=np.float32)
        else:
            rows = np.asarray(list(parameters)[1:37], dtype=np.float32).reshape(4, 9)
        w = np.zeros((4, 10), dtype=np.float32)
        w[:, :9] = rows
        return w


def counts_from_engine(a_book, side, path):
    """parameter_learn_from_edax_protocol.counts (:7-13): ask an external engine for the features of a
    position -- ``<path> -h "<64 chars> <turn>"`` prints a Python list literal -- and prepend the disc
    count.  Works with any engine that implements the switch, e.g.
    ``python -m subproc_b200.edax_engine``.  (``side`` is ignored, as in the reference: the engine sees
    only the position and whose turn it is.)"""
    import subprocess
    from ast import literal_eval
    a_board = board_from_a_book(a_book)
    sfen = a_board.serialize_str()
    out = subprocess.check_output((path + " -h \"%s\"") % sfen, shell=True, universal_newlines=True)
    discs = 64 - a_book['book'].count('-')
    return [discs] + [int(x) for x in literal_eval(out.strip())]


class LearnFromEdaxProtocolProcessParameter(ProgressPositionMovesParameter):
    """parameter_learn_from_edax_protocol.py:16-39: same table, header 3, features from an engine process
    named by ``learn_learn_for_path``"""

    def __init__(self):
        super(LearnFromEdaxProtocolProcessParameter, self).__init__()
        self.conf = None

    def configure(self, conf):
        self.conf = conf

    def header(self):
        return 3

    def hash_from_book(self, a_book, side):
        return ':'.join(str(x) for x in counts_from_engine(a_book, side, self.conf['learn_learn_for_path']))
