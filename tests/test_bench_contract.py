"""bench.py contract checks that need no GPU: the reference arm (the reference's CPU algorithm on the host
cores) prints exactly ONE JSON line with the driver's keys; under torchrun only rank 0 prints and the other
ranks exit 0 without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
        'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e']


def _check(line, n):
    d = json.loads(line)
    for k in KEYS:
        assert k in d, k
    assert d['impl'] == 'reference' and d['n_gpus'] == n and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['value'] > 0 and d['unit'] == 'samples/s' and d['data'] == 'synthetic'
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_single_process():
    out = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--workload', 'cfg1', '--steps', '2',
                          '--warmup', '1'], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_prints_once():
    env = dict(os.environ, OMP_NUM_THREADS='2')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29577', 'bench.py', '--impl', 'reference',
                          '--gpus', '2', '--workload', 'cfg1', '--steps', '1', '--warmup', '0'],
                         cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    _check(lines[0], 2)
