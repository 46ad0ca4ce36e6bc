#!/bin/bash
# timing experiments on the staggered tower (results are wrong under flags): 1 no tap shifts (aligned A), 2 no masks,
# 4 no epilogue math, 8 no weight streaming
for f in 0 1 2 4 8 3 12 15; do
  echo "== AO_TOWER_XFLAGS=$f"
  AO_TOWER_XFLAGS=$f timeout 200 python tools/perf_selfplay.py --rounds 100 --reps 3 --mode ${1:-0} 2>&1 | tail -3
done
