# warp instructions + issue utilisation of one launch of a kernel (regex $1), second launch of the probe
k=${1:-seed_gen}
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:$k -s 1 -c 1 python tools/map_probe.py 100000 1 2>&1 | grep -E "inst_executed|issue_active|time_duration|alu_cycles|seed_|ratio" | head -12
