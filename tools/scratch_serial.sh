#!/bin/bash
for v in default 0 1; do
  if [ "$v" = default ]; then unset BA_CUDA_SERIAL_K2; else export BA_CUDA_SERIAL_K2=$v; fi
  echo "== SERIAL_K2=$v"
  python tools/bench_workloads.py --only cfg1,cfg2,cfg3,cfg4,cfg2x64 --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ',d['workload'],round(d['x_realtime_aggregate'],1),'ms',round(d['ms_per_step'],3),'K1',round(d['channelize_ms_per_step'],3),'K2',round(d['demod_ms_per_step'],3))
"
done
