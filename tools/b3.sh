#!/bin/bash
# tools/b3.sh <label> [env assignments...]: one line with the cfg3 / cfg5 iteration times of the single-pass multinomial kernel
label=$1; shift
for w in cfg3 cfg5; do
  env "$@" python bench.py --workload $w --steps 8 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d['config']['launch']
print('$label $w ms %.3f kernel %.3f NS %s CL %s clusters %s' % (d['ms_per_step'], d['roofline'].get('k_single_ms') or -1, l.get('stages'), l.get('cluster_size'), l.get('clusters')))"
done
