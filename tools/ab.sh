#!/bin/bash
# tools/ab.sh LIB... -- time the playout + greedy kernels with alternative builds of the library
for lib in "$@"; do
  cp "$lib" subproc_b200/libothello_b200.so
  echo "== $lib"
  python bench.py --no-cpu-baseline --steps 30 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('playout', d['value'], d['ms_per_step'])"
  python tools/bench_configs.py --workload selfplay --steps 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('greedy', d['positions_per_s'], d['ms_per_step'])"
done
