#!/bin/bash
# Profiling pass of one round (run under gpurun on one B200):  bash tools/profile_round.sh <tag>
# 1. launch list of a short bench run, 2. ncu --set full of the kernels that can be replayed in place,
# 3. seed_win_kernel under application replay (it rewrites its multi-GB input in place, kernel replay is wrong).
tag=${1:-r1}
out=gpurun_out
cmd="python bench.py --steps 2 --warmup 3 --pairs 100000 --no-cpu-baseline --no-msa"
$cmd > $out/plain_$tag.log 2>&1 || { tail -5 $out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_$tag.csv $cmd > $out/ncu_launch_$tag.log 2>&1
# 70 000 pairs = 280 000 read orientations: one sub-chunk of the LongHit stream, i.e. every seeding launch holds all of them
probe="python tools/map_probe.py 70000 2"
$probe > $out/plain_probe_$tag.log 2>&1 || { tail -5 $out/plain_probe_$tag.log; exit 1; }
# second repetition of the probe: skip the kernels of the first one
ncu --set full --clock-control none --import-source on -k regex:'seed_gen|sw_fill|sw_traceback|pu_count|pu_emit|pair_select' -s 6 -c 6 \
    -o $out/prof_$tag -f $probe > $out/ncu_full_$tag.log 2>&1
small="python tools/map_probe.py 20000 2"
$small > $out/plain_small_$tag.log 2>&1
ncu --replay-mode application --set full --clock-control none --import-source on -k regex:seed_win -s 1 -c 1 \
    -o $out/prof_win_$tag -f $small > $out/ncu_win_$tag.log 2>&1
tail -n 2 $out/ncu_full_$tag.log; tail -n 2 $out/ncu_win_$tag.log
