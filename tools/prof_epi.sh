CMD3="python bench.py --workload cfg3 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD3 > gpurun_out/epi_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_epi_mn' -s 2 -c 1 -f -o gpurun_out/r01c_epi_prof $CMD3 > gpurun_out/epi_ncu.log 2>&1
tail -1 gpurun_out/epi_ncu.log | cut -c1-100
