# ncu captures behind profiles/r02c_*: each only after the same command exited 0 without ncu
set -x
C3="python bench.py --workload cfg3 --n-local 20000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check --secondary ''"
SP="python bench.py --workload spec1 --n-local 50000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check --secondary ''"
eval $C3 > gpurun_out/r02c_cfg3_plain.log 2>&1 && \
eval ncu --set full --clock-control none --import-source on -k regex:k_fused_mn -s 1 -c 1 -f -o gpurun_out/r02c_cfg3_prof $C3 > gpurun_out/r02c_cfg3_ncu.log 2>&1
eval ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_cfg3_n20000.csv $C3 > gpurun_out/r02c_cfg3_ncu2.log 2>&1
eval $SP > gpurun_out/r02c_spec_plain.log 2>&1 && \
eval ncu --set full --clock-control none --import-source on -k regex:"k_spec_fused\|k_spec_grad" -s 2 -c 2 -f -o gpurun_out/r02c_spec1_prof $SP > gpurun_out/r02c_spec_ncu.log 2>&1
eval ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02c_launches_spec1_n50000.csv $SP > gpurun_out/r02c_spec_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
