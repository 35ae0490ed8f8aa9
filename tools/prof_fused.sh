# ncu --set full of the single-pass cluster kernel (cfg2 sample shape, N = 40000 = 21 GB)
CMD="python bench.py --workload cfg2 --n-local 40000 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/fused_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fused_std' -s 1 -c 1 -f -o gpurun_out/r01f_fused_prof $CMD > gpurun_out/fused_ncu.log 2>&1
tail -1 gpurun_out/fused_ncu.log | cut -c1-100
