#!/bin/bash
# tools/ab_learn.sh LIB... -- time othello_learn_accumulate with alternative library builds
for lib in "$@"; do
  cp "$lib" subproc_b200/libothello_b200.so
  python tools/bench_misc.py 2>&1 | grep learn_accumulate | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '%.4g' % d['value'])"
done
