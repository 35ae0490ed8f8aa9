run() { timeout 200 python bench.py --no-e2e --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$*', round(d['ms_per_step'],3), 'fwd', round(r.get('k_fwd_ms',0),3), 'grad', round(r.get('k_grad_ms',0),3), d['final_loss'])"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
run --workload cfg3
run --workload cfg5
run --workload cfg1
