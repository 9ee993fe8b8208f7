#!/bin/bash
# A/B of one environment knob on the default workload, device-resident only: ab_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 5 --no-cpu-baseline --no-other-workloads --verify-stride 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value']), {k: round(v,2) for k,v in d['roofline']['kernels_ms_per_step'].items()}, round(d['e2e']['value']))"
done
