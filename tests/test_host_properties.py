"""Property tests (hypothesis) of the host-side codecs: value-table keys, parameter file bytes, hand
strings, book strings.  No GPU."""
from hypothesis import given, settings, strategies as st

from subproc_b200 import board, paramgen, parameter, value_table

FEATURE_MAX = (64, 33, 4, 8, 4, 8, 8, 16, 4, 12)       # discs, mobility, sizes of the square classes a..h


@settings(max_examples=300, deadline=None)
@given(st.tuples(*[st.integers(0, m) for m in FEATURE_MAX]))
def test_value_table_key_round_trip_and_width(f):
    k = value_table.pack_key(f)
    assert 0 <= k < (1 << 43) and value_table.unpack_key(k) == f
    assert (k >> 36) == f[0]                               # the disc count is the top field (phase filter)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 3), st.lists(st.integers(-127, 127), min_size=36, max_size=36))
def test_parameter_file_round_trip(header, weights):
    params = tuple([header] + weights)
    data = paramgen.encode(params)
    assert len(data) == 38 and data[-1] == 0 and paramgen.decode(data) == params
    table = parameter.ProgressPositionMovesParameter().weights_table(params)
    assert table.shape == (4, 10) and table[:, :9].reshape(-1).tolist() == [float(v) for v in weights]


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 7), st.integers(0, 7), st.sampled_from(['', 'W', 'B', 'WB']))
def test_hand_strings_round_trip(x, y, prefix):
    b = board.Board()
    s = b.handstr_from_coord(x, y)
    assert b.coord_from_handstr(prefix + s) == (x, y) and b.coord_from_handstr(prefix + s.upper()) == (x, y)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 2 ** 64 - 1), st.integers(0, 2 ** 64 - 1))
def test_book_string_round_trip(a, c):
    black, white = a & ~c, c & ~a
    b = board.Board()
    b._black, b._white = black, white
    s = b.serialize_board()
    assert len(s) == 64 and parameter.bits_from_book_string(s) == (black, white)
    d = board.Board()
    d.deserialize(s, 'X', 7)
    assert (d._black, d._white, d.turn, d.nturn) == (black, white, board.White, 7)
