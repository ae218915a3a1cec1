"""oracle/refshim.py -- load the reference's own rules modules (TEST INFRASTRUCTURE).

Gives tests, oracle/make_golden.py and bench.py's CPU-baseline leg the reference's
``board`` / ``parameter`` / ``parameter_progress_position_moves_learn`` modules:

* from ``oracle/_ref`` when oracle/build_ref.py has produced it (the only form that exists
  on the GPU box), else
* transcribed in memory from ``/root/reference`` (this container).

Returns ``None`` when neither exists.  The product package never imports this.
"""
import importlib.util
import os
import sys
import types

from . import build_ref

_NAMES = ("board", "parameter", "parameter_progress_position_moves_learn")
_cache = {}


def available():
    return (os.path.isfile(os.path.join(build_ref.REF_OUT, "board.py"))
            or os.path.isfile(os.path.join(build_ref.REF_SRC, "board.py")))


def source():
    """'oracle/_ref', '/root/reference' or None -- where the modules come from."""
    if os.path.isfile(os.path.join(build_ref.REF_OUT, "board.py")):
        return build_ref.REF_OUT
    if os.path.isfile(os.path.join(build_ref.REF_SRC, "board.py")):
        return build_ref.REF_SRC
    return None


def load():
    """Returns a namespace with .board, .parameter, .ppml (reference modules) or None."""
    if "ns" in _cache:
        return _cache["ns"]
    src = source()
    if src is None:
        return None
    mods = {}
    for name in _NAMES:
        path = os.path.join(src, name + ".py")
        with open(path, "r") as f:
            text = f.read()
        if src == build_ref.REF_SRC:
            text = build_ref.transcribe(name + ".py", text)
        mod = types.ModuleType(name)
        mod.__file__ = path
        # the reference modules import each other by bare name
        sys.modules[name] = mod
        exec(compile(text, path, "exec"), mod.__dict__)
        mods[name] = mod
    ns = types.SimpleNamespace(board=mods["board"], parameter=mods["parameter"],
                               ppml=mods["parameter_progress_position_moves_learn"], source=src)
    _cache["ns"] = ns
    return ns
