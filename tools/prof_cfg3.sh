CMD="python bench.py --workload cfg3 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/r01b_cfg3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01b_launches_cfg3.csv $CMD > gpurun_out/r01b_cfg3_ncu2.log 2>&1
tail -2 gpurun_out/r01b_cfg3_ncu2.log | cut -c1-300
