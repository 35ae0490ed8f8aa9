set -x
CMD3="python bench.py --workload cfg3 --n-local 20000 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD3 > gpurun_out/r01c_cfg3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fwd|k_grad' -s 2 -c 2 -f -o gpurun_out/r01c_cfg3_prof $CMD3 > gpurun_out/r01c_cfg3_ncu.log 2>&1
tail -1 gpurun_out/r01c_cfg3_ncu.log | cut -c1-120
CMD3F="python bench.py --workload cfg3 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD3F > gpurun_out/r01c_cfg3_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01c_launches_cfg3.csv $CMD3F > gpurun_out/r01c_cfg3_ncu2.log 2>&1
CMD2="python bench.py --workload cfg2 --n-local 40000 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD2 > gpurun_out/r01c_cfg2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r01c_launches_cfg2_n40000.csv $CMD2 > gpurun_out/r01c_cfg2_ncu.log 2>&1
tail -1 gpurun_out/r01c_cfg2_ncu.log | cut -c1-120
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
