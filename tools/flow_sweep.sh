run() { timeout 120 python bench.py --workload cfg3 --no-e2e --no-cpu --flow 1 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],2), d['config']['launch'].get('lag_samples'), d['final_loss'])"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -k flow -x -q 2>&1 | tail -3
run --flow-window-mb 32
run --flow-window-mb 64
run --flow-window-mb 96
run --flow-window-mb 32 --flow-debug 2
run --flow-window-mb 64 --flow-debug 2
run --flow-window-mb 32 --flow-debug 6
run --flow-window-mb 64 --flow-debug 6
run --flow-window-mb 128 --flow-debug 6
