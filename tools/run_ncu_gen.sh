# ncu --set full of one launch of the stream kernel (first launch skipped: cold), source view on
k=${1:-seed_gen3}
python tools/map_probe.py 100000 1 > gpurun_out/plain_gen.log 2>&1 || { tail -5 gpurun_out/plain_gen.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_gen -f python tools/map_probe.py 100000 1 > gpurun_out/ncu_gen.log 2>&1
tail -3 gpurun_out/ncu_gen.log
