for w in cfg1 cfg3 cfg2; do python bench.py --workload $w --no-cpu 2>gpurun_out/e2e_$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$w', 'value', round(d['value']), 'e2e', round(e['value']), 'fit_call', e.get('fit_call'))"; tail -2 gpurun_out/e2e_$w.err; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "api or golden or pickle or out_of_core" 2>&1 | tail -3
