"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/othello_b200.h declares, the product never touches oracle/, and the host-only parts of the
Board facade (strings, accessors) behave like the reference (no compute call is made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "othello_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(othello_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = header_symbols()
    for must in ("othello_legal", "othello_flips", "othello_step", "othello_counts", "othello_mask_count",
                 "othello_features", "othello_eval", "othello_playout", "othello_perft",
                 "othello_learn_accumulate", "othello_playout_host", "othello_step_host", "othello_legal_host"):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    from subproc_b200 import build, _lib
    so = build.build()
    assert os.path.isfile(so)
    L = ctypes.CDLL(so)
    for name in header_symbols():
        assert hasattr(L, name), "libothello_b200.so does not export %s" % name
    # the ctypes binding covers the header, no more, no less
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib().othello_abi_version() == _lib.ABI_VERSION
    assert _lib.lib().othello_error_string(-1) == b"invalid argument"


def test_argument_validation_needs_no_gpu():
    from subproc_b200 import _lib
    L = _lib.lib()
    assert L.othello_legal(None, None, None, 5, None) == -1          # OTHELLO_E_INVALID
    assert L.othello_legal(None, None, None, 0, None) == 0           # empty batch is a no-op
    assert L.othello_step(None, None, None, None, None, None, None, None, -1, None) == -1
    assert L.othello_playout(None, None) == -1
    res = ctypes.c_uint64(7)
    assert L.othello_perft(0x0000000810000000, 0x0000001008000000, 1, 0, None, 0, ctypes.byref(res), None) == 0
    assert res.value == 1
    assert L.othello_perft(0, 0, 3, 2, None, 0, ctypes.byref(res), None) == -1
    # round-2 entry points: bad arguments are refused before anything touches a device
    assert L.othello_perft_async(0, 0, 1, 3, 2, 2, None, 0, None, None) == -1            # part >= nparts, no result
    assert L.othello_learn_accumulate(None, None, None, None, None, 5, 5, 120, None, None, None) == -1
    assert L.othello_learn_stats(None, None, None) == -1 and L.othello_learn_refit(None, 1, None, None, None, None, None, None) == -1
    assert L.othello_sort_records(None, None, None, None, 0, 43, None, 0, None) == 0     # nothing to sort
    assert L.othello_sort_records(None, None, None, None, 10, 43, None, 0, None) == -1
    assert L.othello_sort_records(None, None, None, None, 10, 0, None, 0, None) == -1    # key_bits out of range
    assert L.othello_partition_records(None, None, None, None, 4, 300, None, None, 0, None) == -1   # world > 256
    assert L.othello_table_probe(None, -1, None, None, 20, None, 0, None, None) == -1
    assert L.othello_table_apply(None, None, 0, 0.03, None, None, 20, None, None, 0, None, None) == 0
    assert L.othello_table_rehash(None, 10, None, None, 3, None) == -1                   # 10 keys do not fit 8 slots at load 1/2
    assert L.othello_sort_workspace_bytes(1 << 20) >= 256 * 256 * 4 and L.othello_table_workspace_bytes(0) > 0
    assert L.othello_perft_workspace_bytes(2) < L.othello_perft_workspace_bytes(6) <= L.othello_perft_workspace_bytes(40)
    tk = ctypes.c_int64(7)
    assert L.othello_playout_host_async(None, 0, 0, 4, None, None, None, 0, 0, 0, 0, None, -1, None, 120, None, None, None,
                                        None, None, None, None, None, ctypes.byref(tk)) == -1
    assert L.othello_ctx_wait(None, 0) == -1 and L.othello_ctx_set_option(None, 1, 4) == -1
    assert L.othello_board_apply_host(None, 0, 0, 1, 0, None) == -1


def test_product_never_imports_the_oracle_or_the_reference():
    pkg = os.path.join(ROOT, "subproc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "/root/reference" not in text, f


def test_rules_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from subproc_b200 import board
    b = board.Board()
    with pytest.raises(RuntimeError):
        b.puttables(board.Black)
    with pytest.raises(RuntimeError):
        b.put_s('d3')
    from subproc_b200 import batched
    with pytest.raises(RuntimeError):
        batched.BatchedOthello(4)


def test_board_host_side_matches_reference_strings(kat):
    from subproc_b200 import board
    b = board.Board()
    assert (board.Empty, board.Black, board.White) == (0, 1, 2)
    assert board.DIRECS == [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
    assert b.serialize_str() == kat['start_serialize_str']
    assert b.turn == 1 and b.nturn == 0
    for x, y, s in kat['handstr']:
        assert b.handstr_from_coord(x, y) == s
    for c in kat['put_s_cases']:
        assert list(b.coord_from_handstr(c['s'])) == c['coord'], c['s']
    assert [b.string_from_turn(c) for c in (0, 1, 2)] == kat['turn_strings']['string_from_turn']
    assert {s: b.turn_from_string(s) for s in ('O', 'X', '-', '?')} == kat['turn_strings']['turn_from_string']
    assert [b.str_from_turn(c) for c in (0, 1, 2)] == kat['turn_strings']['str_from_turn']
    # deserialize keeps a str nturn as handed over by Redis (parameter.py:7 -> board.py:262)
    d = board.Board()
    d.deserialize(kat['after_d3']['ser_noturn'], kat['after_d3']['ser'][-1], '1')
    assert ("%016x" % d._black, "%016x" % d._white, d.turn, d.nturn) == \
        (kat['deser']['b'], kat['deser']['w'], kat['deser']['turn'], kat['deser']['nturn'])
    assert d.serialize_str() == kat['after_d3']['ser']
    assert board.clone_board(d.board) == d.board and d.serialize_tuple() == (d.board, d.turn)
    assert d.get(3, 2) == board.Black and d.get(0, 0) == board.Empty
    with pytest.raises(IndexError):
        d.get(8, 0)
    assert board.is_within_board(7, 7) and not board.is_within_board(8, 0) and not board.is_within_board(0, -1)
