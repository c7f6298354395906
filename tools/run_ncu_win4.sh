# ncu --set full of seed_win_kernel on cfg4 (application replay: the kernel sweeps its input in place)
small="python tools/map_probe.py 20000 2 4"
$small > gpurun_out/plain_win4.log 2>&1 || { tail -5 gpurun_out/plain_win4.log; exit 1; }
ncu --replay-mode application --set full --clock-control none --import-source on -k regex:seed_win -s 1 -c 1 -o gpurun_out/prof_win_cfg4 -f $small > gpurun_out/ncu_win4.log 2>&1
tail -3 gpurun_out/ncu_win4.log
