#!/bin/bash
# tools/b2.sh <label> [env assignments...]: iteration time of the default workload (cfg2, standard model, single-pass cluster kernel)
label=$1; shift
env "$@" python bench.py --workload cfg2 --secondary "" --steps 8 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
l=d['config']['launch']; print('$label cfg2 ms %.3f kernel %.3f CL %s clusters %s stages %s' % (d['ms_per_step'], r.get('k_single_ms') or -1, l.get('cluster_size'), l.get('clusters'), l.get('stages')))"
