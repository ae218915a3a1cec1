"""oracle/build_ref.py -- recipe that makes the reference's own rules code runnable here.

TEST INFRASTRUCTURE.  The reference (ysnrkdm/subproc) is Python 2; this image has only
CPython 3.12 and no 2to3.  The three files on the hot path need exactly two mechanical
source substitutions to run under Python 3:

    board.py:92   ``print q``   -> ``print(q)``      (debug helper show_mask)
    board.py:257  ``i / 8``     -> ``i // 8``        (py2 integer division in deserialize)

``parameter.py`` and ``parameter_progress_position_moves_learn.py`` parse unmodified.

This script reads the sources WHERE THEY LIE (``/root/reference``), applies the two
substitutions and writes the result into the git-ignored ``oracle/_ref/`` -- the analogue,
for an interpreted reference, of compiling a C reference into ``oracle/_ref/*.so``.  Nothing
from the reference is committed.  ``oracle/_ref`` travels to the GPU box with the gpurun
snapshot so the CPU baseline there times the reference's own code.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SUBPROC_REFERENCE", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")

FILES = ("board.py", "parameter.py", "parameter_progress_position_moves_learn.py")

SUBSTITUTIONS = {
    "board.py": (
        ("        print q\n", "        print(q)\n"),
        ("self.set(cell, i % 8, i / 8)", "self.set(cell, i % 8, i // 8)"),
    ),
}


def transcribe(name, text):
    for old, new in SUBSTITUTIONS.get(name, ()):
        if text.count(old) != 1:
            raise RuntimeError("%s: expected exactly one occurrence of %r" % (name, old))
        text = text.replace(old, new)
    return text


def build(src=REF_SRC, out=REF_OUT):
    """Returns True when oracle/_ref was (re)built, False when the reference is absent."""
    if not os.path.isfile(os.path.join(src, "board.py")):
        return False
    os.makedirs(out, exist_ok=True)
    for name in FILES:
        with open(os.path.join(src, name), "r") as f:
            text = f.read()
        with open(os.path.join(out, name), "w") as f:
            f.write(transcribe(name, text))
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref built" if ok else "reference not found at %s; nothing built" % REF_SRC)
    sys.exit(0)
