# one-GPU round check: parity tests, smoke, every BASELINE config, the reference arm
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r01f_bench_cfg2.json 2> gpurun_out/r01f_bench_cfg2.err
for w in cfg1 cfg3 cfg4 cfg5; do python bench.py --workload $w > gpurun_out/r01f_bench_$w.json 2> gpurun_out/r01f_bench_$w.err; done
python bench.py --impl reference > gpurun_out/r01f_ref_cfg2.json 2> gpurun_out/r01f_ref_cfg2.err
python bench.py --impl reference --workload cfg3 > gpurun_out/r01f_ref_cfg3.json 2> gpurun_out/r01f_ref_cfg3.err
for f in gpurun_out/r01f_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get('roofline') or {}
e=d.get('e2e') or {}
print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', r.get('frac'), 'it_frac', r.get('iteration_frac_of_2pass_roofline'), 'e2e', e.get('value'), 'clocks', d.get('clocks'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
done
