"""oracle/make_golden.py -- generate tests/golden/*.json.gz FROM THE REFERENCE ITSELF.

TEST INFRASTRUCTURE.  Run in a container that has /root/reference:

    python -m oracle.make_golden

Every value written comes out of the reference's own ``board.py`` /
``parameter_progress_position_moves_learn.py`` (loaded through oracle/refshim.py, i.e. with the
two py2->py3 substitutions of oracle/build_ref.py and nothing else).  The game loop of
``game_runner.py`` cannot be imported (py2 prints, needs two engine binaries), so the loop of
``play_a_game``/``go_for`` (game_runner.py:133-184) is restated here in a few lines of Python that
drive the reference Board; the engines are stand-ins that answer like an Edax-protocol engine
playing uniformly at random or greedily on the linear form.  The counter-based RNG is the
build's own (DESIGN.md "RNG"); it is restated here in pure Python so the fixtures pin it too.

The fixtures are the pins for oracle/othello_oracle.c (tests/test_oracle_golden.py) and, on the
GPU box where /root/reference does not exist, for the CUDA path directly (tests/test_gpu_*.py).
"""
import gzip
import json
import os
import sys

import numpy as np

from . import refshim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
M32 = 0xFFFFFFFF

DEFAULT_ROWS = None  # filled from the reference's default_value()


# ---- RNG spec, pure Python ------------------------------------------------------------
def fmix32(h):
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def rng_key(seed, gid):
    h = fmix32((seed & M32) ^ 0x9E3779B9)
    h = fmix32(h ^ ((seed >> 32) & M32))
    h = fmix32(h ^ (gid & M32))
    h = fmix32(h ^ ((gid >> 32) & M32))
    return h


def rng_draw(key, ply, stream):
    return fmix32((key + ply * 0x9E3779B9 + stream * 0x632BE5AB) & M32)


def below(r, n):
    return (r * n) >> 32


# ---- helpers on the reference Board ----------------------------------------------------
def bits(rb, B):
    black = white = 0
    for y in range(8):
        for x in range(8):
            c = B.get(x, y)
            if c == rb.Black:
                black |= 1 << (x + 8 * y)
            elif c == rb.White:
                white |= 1 << (x + 8 * y)
    return black, white


def mask_of(coords):
    m = 0
    for (x, y) in coords:
        m |= 1 << (x + 8 * y)
    return m


def book_of(B):
    return {'book': B.serialize_board(), 'whosturn': B.serialize_turn(), 'turn': B.nturn}


def hx(v):
    return "%016x" % v


def clone(rb, B):
    C = rb.Board()
    C.board = rb.clone_board(B.board)
    C.turn = B.turn
    C.nturn = B.nturn
    return C


def ref_eval(ns, B, side_colour, rows):
    """numpy fp64 dot of counts() with the phase row."""
    side = 'O' if side_colour == ns.board.Black else 'X'
    f = ns.ppml.counts(book_of(B), side)
    discs = f[0]
    row = 0 if discs <= 16 else 1 if discs <= 32 else 2 if discs <= 48 else 3
    return float(np.dot(np.asarray(rows[row], dtype=np.float64), np.asarray(f[1:], dtype=np.float64)))


def engine_answer(ns, B, puttables, greedy, r1, rows):
    """what Player.go would return (game_runner.py:19-33) for a random / greedy engine."""
    rb = ns.board
    if len(puttables) == 0:
        return 'ps'
    if not greedy:
        x, y = puttables[below(r1, len(puttables))]
        return B.handstr_from_coord(x, y)
    best, best_v = None, None
    for (x, y) in puttables:
        C = clone(rb, B)
        C.put_s(C.handstr_from_coord(x, y))
        v = ref_eval(ns, C, B.turn, rows)
        if best is None or v > best_v:
            best, best_v = (x, y), v
    return B.handstr_from_coord(*best)


def play_game(ns, seed, gid, policy, random_plies, n_rand_black, n_rand_white, rows, record_features,
              policy_white=None, rows_white=None):
    """GameRunner.play_a_game / play_a_turn / go_for (game_runner.py:133-184) on the reference Board."""
    rb = ns.board
    B = rb.Board()
    rest = {rb.Black: min(n_rand_black, 10), rb.White: min(n_rand_white, 10)}   # game_runner.py:118-119
    key = rng_key(seed, gid)
    plies = []
    positions = []

    def snap():
        black, white = bits(rb, B)
        rec = {'b': hx(black), 'w': hx(white), 'turn': B.turn, 'nturn': B.nturn,
               'legal_b': hx(mask_of(B.puttables(rb.Black))), 'legal_w': hx(mask_of(B.puttables(rb.White))),
               'over': bool(B.is_game_over()), 'nb': B.n_black(), 'nw': B.n_white(), 'ne': B.n_empty(),
               'ser': B.serialize_str()}
        if record_features:
            rec['feat_O'] = list(ns.ppml.counts(book_of(B), 'O'))
            rec['feat_X'] = list(ns.ppml.counts(book_of(B), 'X'))
        positions.append(rec)

    snap()
    t = 0
    over = B.is_game_over()
    while not over:
        r0, r1 = rng_draw(key, t, 0), rng_draw(key, t, 1)
        puttables = B.puttables(B.turn)
        hand = None
        if rest[B.turn] > 0 and below(r0, rest[B.turn]) == 0:                     # game_runner.py:134-135
            if len(puttables) > 0:
                x, y = puttables[below(r1, len(puttables))]
                hand = B.handstr_from_coord(x, y)
                rest[B.turn] -= 1
        if hand is None:
            # proc_black and proc_white may be different engines (game_runner.py:107-123)
            pol = policy if (B.turn == rb.Black or policy_white is None) else policy_white
            rws = rows if (B.turn == rb.Black or rows_white is None) else rows_white
            hand = engine_answer(ns, B, puttables, pol == 1 and t >= random_plies, r1, rws).lower()
        before = bits(rb, B)
        mover = B.turn
        ret = B.put_s(hand)
        after = bits(rb, B)
        own_before = before[0] if mover == rb.Black else before[1]
        own_after = after[0] if mover == rb.Black else after[1]
        if hand == 'ps':
            move, flips = 64, 0
        else:
            x, y = B.coord_from_handstr(hand)
            move = x + 8 * y
            flips = (own_after ^ own_before) & ~(1 << move)
        plies.append({'hand': hand, 'move': move, 'ret': ret, 'flips': hx(flips)})
        snap()
        t += 1
        over = B.is_game_over()
    out = {'seed': seed, 'gid': gid, 'policy': policy, 'random_plies': random_plies,
           'n_rand_black': n_rand_black, 'n_rand_white': n_rand_white,
           'plies': plies, 'positions': positions}
    if policy_white is not None:
        out['policy_white'] = policy_white
    if rows_white is not None:
        out['rows_white'] = rows_white
    return out


def perft(rb, B, depth):
    if depth == 0:
        return 1
    moves = B.puttables(B.turn)
    if len(moves) == 0:
        if B.is_game_over():
            return 1
        C = clone(rb, B)
        C.put_s('ps')
        return perft(rb, C, depth - 1)
    total = 0
    for (x, y) in moves:
        C = clone(rb, B)
        C.put_s(C.handstr_from_coord(x, y))
        total += perft(rb, C, depth - 1)
    return total


def make_kat(ns):
    rb = ns.board
    B = rb.Board()
    black, white = bits(rb, B)
    kat = {
        'start_black': hx(black), 'start_white': hx(white),
        'start_puttables_black': [list(c) for c in B.puttables(rb.Black)],
        'start_puttables_white': [list(c) for c in B.puttables(rb.White)],
        'start_serialize_str': B.serialize_str(),
        'start_str': str(B),
        'start_counts_O': list(ns.ppml.counts(book_of(B), 'O')),
        'start_counts_X': list(ns.ppml.counts(book_of(B), 'X')),
        'perft': [perft(rb, rb.Board(), d) for d in range(1, 7)],
        'default_value': ns.ppml.ProgressPositionMovesParameter().default_value(),
        'header': ns.ppml.ProgressPositionMovesParameter().header(),
    }
    # put_s on strings, from the start position (board.py:176-209)
    cases = []
    for s in ['d3', 'D3', 'Bd3', 'Wd3', 'WBd3', 'c4', 'f5', 'e6', 'a1', 'd4', 'e3', 'ps', 'PS', 'Ps', 'pS',
              '', 'zz', 'a0', 'i1', 'i9', 'a9', 'h8', 'd3e6', '3d', 'pass', ' d3 ', 'x']:
        C = rb.Board()
        try:
            ret = C.put_s(s)
            err = None
        except Exception as e:                                   # IndexError for x > 7 or rank 9
            ret, err = None, type(e).__name__
        b2, w2 = bits(rb, C)
        cases.append({'s': s, 'ret': ret, 'err': err, 'b': hx(b2), 'w': hx(w2), 'turn': C.turn, 'nturn': C.nturn,
                      'coord': list(rb.Board().coord_from_handstr(s))})
    kat['put_s_cases'] = cases
    kat['handstr'] = [[x, y, rb.Board().handstr_from_coord(x, y)] for y in range(8) for x in range(8)]
    # after d3
    C = rb.Board()
    r = C.put_s('d3')
    kat['after_d3'] = {'ret': r, 'nb': C.n_black(), 'nw': C.n_white(), 'turn': C.turn, 'nturn': C.nturn,
                       'str': str(C), 'ser': C.serialize_str(), 'ser_noturn': C.serialize_str(False)}
    # deserialize round trip incl. the string-typed nturn that Redis would hand back (parameter.py:7)
    D = rb.Board()
    D.deserialize(C.serialize_board(), C.serialize_turn(), '1')
    kat['deser'] = {'b': hx(bits(rb, D)[0]), 'w': hx(bits(rb, D)[1]), 'turn': D.turn, 'nturn': D.nturn}
    kat['turn_strings'] = {'string_from_turn': [B.string_from_turn(c) for c in (0, 1, 2)],
                           'turn_from_string': {s: B.turn_from_string(s) for s in ('O', 'X', '-', '?')},
                           'str_from_turn': [B.str_from_turn(c) for c in (0, 1, 2)]}
    return kat


def make_probe(ns, games, n_positions=160):
    """put() / mask_count() of every square for BOTH colours on positions sampled from the games."""
    rb = ns.board
    rng = np.random.RandomState(7)
    pool = [p for g in games for p in g['positions']]
    picks = rng.choice(len(pool), size=min(n_positions, len(pool)), replace=False)
    out = []
    for i in picks:
        p = pool[int(i)]
        B = rb.Board()
        B.deserialize(p['ser'][:64], p['ser'][65], p['nturn'])
        rec = {'b': p['b'], 'w': p['w'], 'put': []}
        for piece in (rb.Black, rb.White):
            rets, flips = [], []
            for s in range(64):
                C = clone(rb, B)
                before = bits(rb, C)
                ret = C.put(piece, s & 7, s >> 3)
                after = bits(rb, C)
                own_b = before[0] if piece == rb.Black else before[1]
                own_a = after[0] if piece == rb.Black else after[1]
                rets.append(ret)
                flips.append(hx((own_a ^ own_b) & ~(1 << s)))
            rec['put'].append({'piece': piece, 'ret': rets, 'flips': flips})
        m = int(rng.randint(0, 2 ** 32)) | (int(rng.randint(0, 2 ** 32)) << 32)
        rec['mask'] = hx(m)
        rec['mask_count'] = [B.mask_count(rb.Black, m), B.mask_count(rb.White, m)]
        # put_s of every move code for the side to move (illegal ones must return -1 and change nothing)
        ps = []
        for s in range(64):
            C = clone(rb, B)
            ps.append(C.put_s(C.handstr_from_coord(s & 7, s >> 3)))
        rec['put_s'] = ps
        rec['turn'] = p['turn']
        out.append(rec)
    return out


def make_paramgen():
    """paramgen.write_data (paramgen.py:5-19) run on a few parameter tuples.  paramgen.py needs its
    py2 print statements parenthesised and a stub for its ``import config``; nothing else changes."""
    import re
    import tempfile
    import types
    from . import build_ref
    path = os.path.join(build_ref.REF_SRC, "paramgen.py")
    text = open(path).read()
    text = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", text, flags=re.M)
    sys.modules.setdefault("config", types.ModuleType("config"))
    mod = types.ModuleType("paramgen_ref")
    exec(compile(text, path, "exec"), mod.__dict__)
    cases = []
    rng = np.random.RandomState(11)
    params = [
        (2, 100, 99, -1, -1, -1, -1, 3, 8, 20, 75, 99, 2, -5, 7, 6, 4, 5, 5, 25, 99, 2, -5, -7, -6, 4, 5, 5,
         1, 100, 50, 30, 30, 30, 30, 30, 30),
        tuple([3] + [int(v) for v in rng.randint(-127, 128, size=36)]),
        tuple([2] + [-127, 127, 0, -1, 1] + [0] * 31),
    ]
    for p in params:
        with tempfile.NamedTemporaryFile(delete=False) as f:
            name = f.name
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            mod.write_data(name, p)
        data = open(name, "rb").read()
        os.unlink(name)
        cases.append({'params': list(p), 'bytes': data.hex()})
    return cases


def reference_update_rule(ns):
    """The reference's own value-table update, EXECUTED from its source text: the two methods
    ``__update_state_for_a_book`` / ``__update_state_map`` (progress_position_moves_learn.py:37-62) and the
    constants of ``__init__`` (:19-25) are cut out of /root/reference/progress_position_moves_learn.py and
    compiled as they stand (the module itself cannot be imported: py2 print statements, redis / pyres /
    sklearn-era imports).  Substitutions: the py2 debug statement ``print book[0]`` is dropped; nothing else.
    ``self._param_store()`` -- Redis in the reference -- is a dict with the three calls the text makes
    (exists / set / get); Redis hands strings back, so ``get`` returns ``repr`` text like redis-py would.
    Returns an object whose ``update(book_id, book)`` is the reference's ``__update_state_for_a_book``."""
    import re
    import textwrap
    from . import build_ref
    path = os.path.join(build_ref.REF_SRC, "progress_position_moves_learn.py")
    text = open(path).read()
    lines = text.split("\n")
    start = next(i for i, ln in enumerate(lines) if re.match(r"\s+def __update_state_for_a_book\(", ln))
    end = next(i for i, ln in enumerate(lines) if i > start and ln.strip() == "# Learning related functions")
    body = [ln for ln in lines[start:end] if ln.strip() != "print book[0]       # terminal book"]
    assert len(body) == end - start - 1, "the py2 debug print was expected exactly once"
    consts = dict(re.findall(r"self\.(a|l) = ([0-9.]+)", text))
    src = "class ReferenceRule(object):\n" + "\n".join(body) + "\n"
    src += textwrap.dedent("""
        def update(self, book_id, book):
            return self._ReferenceRule__update_state_for_a_book(book_id, book)
    """).replace("\n", "\n    ")

    class Store(object):
        def __init__(self):
            self.d = {}

        def exists(self, key):
            return tuple(key) in self.d

        def set(self, key, value):
            self.d[tuple(key)] = repr(value)         # what lands in Redis (py2.7 / py3 repr round-trips floats)

        def get(self, key):
            return self.d[tuple(key)]

    scope = {"board_from_a_book": ns.parameter.board_from_a_book}
    exec(compile(src, path + ":37-62", "exec"), scope)
    rule = scope["ReferenceRule"]()
    rule.a, rule.l = float(consts["a"]), float(consts["l"])
    rule.parameter = ns.ppml.ProgressPositionMovesParameter()
    store = Store()
    rule._param_store = lambda: store
    rule.store = store
    return rule


def make_value_table(ns, games, batches=((0, 12), (12, 24))):
    """the value table the reference builds from the first 24 golden games (seed 0, uniform random, game ids
    0..23), in two batches: books sorted by turn and reversed, terminal record first (replearn.py:37-38)"""
    rule = reference_update_rule(ns)
    out = {"a": rule.a, "l": rule.l, "seed": 0, "batches": [list(b) for b in batches], "after": []}
    for lo, hi in batches:
        for gid in range(lo, hi):
            g = games[gid]
            assert g['seed'] == 0 and g['gid'] == gid and g['policy'] == 0
            recs = [{'book': p['ser'][:64], 'whosturn': p['ser'][65], 'turn': str(p['nturn']), 'end': p['over']}
                    for p in g['positions']]
            book = list(reversed(sorted(recs, key=lambda x: int(x['turn']))))
            assert book[0]['end']
            rule.update(gid, book)
        out["after"].append({k[2]: float(v).hex() for k, v in rule.store.d.items()})
    return out


def dump(name, obj):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name)
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(',', ':'), sort_keys=True).encode())
    print("wrote %s (%d bytes)" % (path, os.path.getsize(path)))


def main():
    ns = refshim.load()
    if ns is None:
        print("reference not available; cannot generate golden vectors")
        return 1
    rows = ns.ppml.ProgressPositionMovesParameter().default_value()
    dump("kat.json.gz", make_kat(ns))
    games = []
    for gid in range(24):                                       # config 1 / 3 style: uniform random
        games.append(play_game(ns, 0, gid, 0, 0, 0, 0, rows, record_features=(gid < 8)))
    for gid in range(8):                                        # go_for substitution over a random engine
        games.append(play_game(ns, 1, gid, 0, 0, 3, 10, rows, record_features=False))
    for gid in range(6):                                        # config 4: R random plies, then greedy
        games.append(play_game(ns, 2, gid, 1, 10, 0, 0, rows, record_features=False))
    for gid in range(4):                                        # greedy engine + reference substitution rule
        games.append(play_game(ns, 3, gid, 1, 0, 10, 2, rows, record_features=False))
    other = [[3, 80, 40, -20, 5, 5, 1, 1, 2], [10, 60, 30, -10, 4, 6, 2, 2, 1],
             [20, 50, 20, -5, 3, 7, 3, 3, 3], [64, 10, 10, 10, 10, 10, 10, 10, 10]]
    for gid in range(3):                                        # different engines per colour: greedy vs random
        games.append(play_game(ns, 4, gid, 1, 2, 0, 0, rows, False, policy_white=0))
    for gid in range(3):                                        # two greedy engines with different parameter sets
        games.append(play_game(ns, 5, gid, 1, 4, 1, 1, rows, False, policy_white=1, rows_white=other))
    dump("games.json.gz", games)
    dump("probe.json.gz", make_probe(ns, games))
    dump("paramgen.json.gz", make_paramgen())
    dump("value_table.json.gz", make_value_table(ns, games))
    n_pass = sum(1 for g in games for p in g['plies'] if p['move'] == 64)
    print("games=%d plies=%d passes=%d" % (len(games), sum(len(g['plies']) for g in games), n_pass))
    return 0


if __name__ == "__main__":
    sys.exit(main())
