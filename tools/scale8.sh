set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --workload cfg2 > gpurun_out/r01c_scale8_cfg2.json 2> gpurun_out/r01c_scale8_cfg2.err
$TR bench.py --gpus 8 --workload cfg3 > gpurun_out/r01c_scale8_cfg3.json 2> gpurun_out/r01c_scale8_cfg3.err
$TR bench.py --gpus 8 --workload cfg5 --no-e2e > gpurun_out/r01c_scale8_cfg5.json 2> gpurun_out/r01c_scale8_cfg5.err
$TR bench.py --gpus 8 --impl reference > gpurun_out/r01c_scale8_ref.json 2> gpurun_out/r01c_scale8_ref.err
for f in gpurun_out/r01c_scale8_*.json; do tail -1 $f | cut -c1-420; done
