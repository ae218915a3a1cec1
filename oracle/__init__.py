"""oracle/ -- CPU restatement of the reference's rules and learner arithmetic.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product package
``subproc_b200`` never does (tests/test_boundary.py greps for it).

``oracle.othello_oracle.c``  plain-C mailbox restatement (every function cites board.py etc.)
``oracle.lib``               ctypes + numpy wrapper around it (builds it with gcc on demand)
``oracle.refshim``           loader for the reference's own Python modules (py3 transcription)
``oracle.build_ref``         recipe that writes that transcription into git-ignored oracle/_ref/
``oracle.make_golden``       generates tests/golden/*.json from the reference itself

Parity status: pinned against the reference's own board.py (see othello_oracle.c header).
"""
