#!/bin/bash
# tools/bspec.sh <label> [env assignments...]: one line with the iteration time of the spectral bench workload (spec1)
label=$1; shift
env "$@" python bench.py --workload spec1 --secondary "" --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d['config']['launch']; r=d['roofline']
print('$label spec1 ms %.3f kernel %s single %s fwd %s grad %s stages %s frac %.3f' % (d['ms_per_step'], r['kernel'], r.get('k_single_ms'), r.get('k_fwd_ms'), r.get('k_grad_ms'), l.get('stages'), r['iteration_frac_of_2pass_roofline']))"
