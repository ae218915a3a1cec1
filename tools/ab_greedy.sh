#!/bin/bash
# tools/ab_greedy.sh LIB... -- time the greedy self-play kernel with alternative library builds
for lib in "$@"; do
  cp "$lib" subproc_b200/libothello_b200.so
  python tools/bench_configs.py --workload selfplay --steps 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'greedy', '%.4g' % d['positions_per_s'], d['ms_per_step'])"
done
