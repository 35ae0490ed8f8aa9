set -x
CMD="python bench.py --workload cfg3 --n-local 20000 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/r01b_cfg3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fwd|k_grad' -s 2 -c 2 -f -o gpurun_out/r01b_cfg3_prof $CMD > gpurun_out/r01b_cfg3_ncu.log 2>&1
tail -3 gpurun_out/r01b_cfg3_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01b_launches_cfg3_n20000.csv $CMD > gpurun_out/r01b_cfg3_ncu2.log 2>&1
tail -2 gpurun_out/r01b_cfg3_ncu2.log
