# ncu capture of the single-pass spectral kernel (after the same command exited 0 without ncu)
tag=${1:-r02d}
SP="python bench.py --workload spec1 --n-local 50000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-check --secondary ''"
eval $SP > gpurun_out/${tag}_spec_plain.log 2>&1 && \
eval ncu --set full --clock-control none --import-source on -k regex:k_spec_single -s 1 -c 1 -f -o gpurun_out/${tag}_spec_single_prof $SP > gpurun_out/${tag}_spec_ncu.log 2>&1
eval ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches_spec1_single_n50000.csv $SP > gpurun_out/${tag}_spec_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
