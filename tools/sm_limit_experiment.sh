#!/bin/bash
# What a cluster shape that strands SMs would start from (clusters of 4 leave 16 of the 148 SMs idle):
# the persistent self-play kernel on 148 / 140 / 132 SMs (probe build, AO_SM_LIMIT), with SM clock and board power
# sampled during each run.  The 4096 games stay the same, so fewer SMs = more games per CTA.
export AO_USE_PROBE_LIB=1
for n in 148 140 132; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 250 > /tmp/smi_$n.csv &
  SMI=$!
  AO_SM_LIMIT=$n python tools/perf_leg.py selfplay9 1600 | tail -1
  kill $SMI
  python - <<PY
import statistics
rows=[l.split(',') for l in open('/tmp/smi_$n.csv') if l.strip()]
rows=rows[len(rows)//3:]
print("  SMs $n: median SM clock %.0f MHz, median board power %.0f W over the last two thirds of the run" % (statistics.median(float(r[0]) for r in rows), statistics.median(float(r[1]) for r in rows)))
PY
done
