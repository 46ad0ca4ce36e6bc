#!/bin/bash
# Round-2 profiling session (run under gpurun, one GPU): launch list of bench.py's timed region + one `ncu --set full`
# capture per hot kernel and configuration.  Outputs under gpurun_out/; summaries are copied to profiles/ by hand.
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 400 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-legs --no-cpu-baseline --no-full-episodes --no-e2e > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:tower_stag -s 40 -c 1 -f -o $O/r02_tower_stag9 python tools/perf_leg.py selfplay9 48 > $O/r02_ncu_stag9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tower_stag -s 40 -c 1 -f -o $O/r02_tower_stag15 python tools/perf_leg.py selfplay15 48 > $O/r02_ncu_stag15.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tower_kernel -s 40 -c 1 -f -o $O/r02_tower_x3 python tools/perf_leg.py trained9 48 > $O/r02_ncu_x3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_step -s 60 -c 1 -f -o $O/r02_tree_step_arena python tools/perf_leg.py arena 48 > $O/r02_ncu_tree_arena.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 90 --csv --log-file $O/r02_launches_arena.csv python tools/perf_leg.py arena 80 > $O/r02_launches_arena.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_search -c 1 -f -o $O/r02_rollout_search python tools/perf_leg.py rollout > $O/r02_ncu_rollout.log 2>&1
for leg in selfplay9 selfplay15 trained9 arena rollout; do python tools/perf_leg.py $leg 400; done > $O/r02_perf_legs.log 2>&1
tail -8 $O/r02_perf_legs.log
# ---- second session of round 2: the cluster-of-four kernel (single game), probes
ncu --set full --clock-control none --import-source on -k regex:tower_solo -s 3 -c 1 -f -o $O/r3_tower_solo9 python tools/perf_single_game.py 400 > $O/r3_ncu_solo.log 2>&1
python tools/probe_single_phases.py 400 > $O/r3_phases.log 2>&1; AO_NO_SOLO=1 python tools/probe_single_phases.py 400 >> $O/r3_phases.log 2>&1
python tools/probe_mma_small_n.py > $O/r3_mma_small_n.log 2>&1; python tools/probe_mma_small_n.py sw128 >> $O/r3_mma_small_n.log 2>&1
bash tools/sm_limit_experiment.sh > $O/r3_sm_limit.log 2>&1
python tools/profile_train_step.py graph > $O/r3_train_profile.log 2>&1; python tools/perf_train_step.py >> $O/r3_train_profile.log 2>&1
for a in "400" "40" "400 9 trained" "40 9 trained" "400 15"; do python tools/perf_single_game.py $a | tail -1; done > $O/r3_single_game.log 2>&1
