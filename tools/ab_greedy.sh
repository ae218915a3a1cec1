#!/bin/bash
# tools/ab_greedy.sh LIB... -- time the greedy kernel (2^16 and 2^19 games) with alternative library builds
for lib in "$@"; do
  cp "$lib" subproc_b200/libothello_b200.so
  echo "== $lib"
  python tools/bench_learn.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('greedy 2^16 ms', d['greedy_playout_ms'])"
  python tools/bench_configs.py --workload selfplay --steps 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('greedy 2^19', '%.4g' % d['positions_per_s'], d['ms_per_step'])"
done
