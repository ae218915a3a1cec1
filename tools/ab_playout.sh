#!/bin/bash
# tools/ab_playout.sh LIB... -- time only the random playout kernel with alternative library builds
for lib in "$@"; do
  cp "$lib" subproc_b200/libothello_b200.so
  python bench.py --no-cpu-baseline --steps 40 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'playout', '%.4g' % d['value'], d['ms_per_step'])"
done
