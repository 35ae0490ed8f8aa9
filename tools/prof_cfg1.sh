CMD="python bench.py --workload cfg1 --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r01c_cfg1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01c_launches_cfg1.csv $CMD > gpurun_out/r01c_cfg1_ncu.log 2>&1
tail -1 gpurun_out/r01c_cfg1_ncu.log | cut -c1-200
