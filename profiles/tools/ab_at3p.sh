#!/bin/bash
# A/B of differently built libraries (ATDE_LIB) on a quarter-size ATRAC3plus batch, device-resident only.
for lib in "$@"; do
  ATDE_LIB=$PWD/atracdenc_b200/$lib python bench.py --workload atrac3plus_stereo --streams 512 --frames 489 --steps 3 --no-cpu-baseline --no-other-workloads --verify-stride 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), {k: round(v,2) for k,v in d['roofline']['kernels_ms_per_step'].items()}, round(d['e2e']['value']))"
done
