#!/bin/bash
# timing / power experiments on the staggered tower (results are wrong under flags): 1 no tap shifts (aligned A),
# 4 no epilogue at all, 8 no weight streaming, 16 no fp32 stash (tcgen05.st), 32 no operand stores (st.shared),
# 64 no accumulator loads (tcgen05.ld)
for f in ${@:-0 16 32 64 48 112 4 8}; do
  echo "== AO_TOWER_XFLAGS=$f"
  AO_USE_PROBE_LIB=1 AO_TOWER_XFLAGS=$f timeout 200 python tools/perf_selfplay.py --rounds 100 --reps 6 2>&1 | tail -3
done
