#!/bin/bash
# end-to-end frames/s of the default workload against the host path's chunk size (MiB of PCM per chunk)
for mib in "$@"; do
  ATDE_CHUNK_MIB=$mib python bench.py --steps 5 --no-cpu-baseline --no-other-workloads --verify-stride 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk MiB $mib', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), 'i16', round(d['e2e']['i16']['value']), round(d['e2e']['i16']['ms_per_step'],1))"
done
