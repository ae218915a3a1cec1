"""Game recorders with the reference's interface (game_recorder.py:9-79), for callers that want the
per-position records of a game.  The Redis / DynamoDB recorders of the reference are network I/O and
out of scope (as is the flat-file recorder: its text format is ``books.flatfile_text``);
``MemoryRecorder`` keeps the RedisRecorder's book dicts in memory."""
from abc import ABCMeta, abstractmethod


class GameRecorder(object, metaclass=ABCMeta):          # game_recorder.py:9-38
    @abstractmethod
    def __enter__(self):
        pass

    @abstractmethod
    def __exit__(self, exc_type, exc_val, exc_tb):
        pass

    @abstractmethod
    def graceful_exit(self):
        pass

    @abstractmethod
    def configure(self, title, meta, config_dict):
        pass

    @abstractmethod
    def add(self, serialize_data):
        pass

    @abstractmethod
    def store(self):
        pass

    @abstractmethod
    def add_meta(self, meta_dict):
        pass


class MemoryRecorder(GameRecorder):
    """RedisRecorder's records (game_recorder.py:107-114) kept in a list instead of Redis hashes"""

    def __init__(self):
        self.lines = []
        self.meta = {'meta': 'meta'}
        self.stored = []

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        return False

    def graceful_exit(self):
        pass

    def configure(self, title, meta, config_dict):
        pass

    def add(self, game_board):
        self.lines.append({'book': game_board.serialize_board(), 'whosturn': game_board.serialize_turn(),
                           'turn': game_board.nturn, 'end': game_board.is_game_over()})

    def store(self):
        self.stored.append((list(self.lines), dict(self.meta)))

    def add_meta(self, meta_dict):
        self.meta.update(meta_dict)
