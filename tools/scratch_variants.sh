#!/bin/bash
# scratch: time cfg2 / cfg4 / cfg1 with alternative builds of demod.cu (BA_CUDA_LIB)
for v in base "$@"; do
  if [ "$v" = base ]; then unset BA_CUDA_LIB; else export BA_CUDA_LIB=$PWD/boondock_airband_b200/csrc/build/variants/libba_cuda_$v.so; fi
  echo "== $v"
  python tools/bench_workloads.py --only cfg1,cfg2,cfg3,cfg4,cfg2x64 --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ',d['workload'],round(d['x_realtime_aggregate'],1),'ms',round(d['ms_per_step'],3),'K1',round(d['channelize_ms_per_step'],3),'K2',round(d['demod_ms_per_step'],3))
"
done
for v in base "$@"; do
  if [ "$v" = base ]; then unset BA_CUDA_LIB; else export BA_CUDA_LIB=$PWD/boondock_airband_b200/csrc/build/variants/libba_cuda_$v.so; fi
  echo "== cfg5 $v"
  python bench.py --steps 8 --warmup 3 --no-cpu --no-workloads --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value ms',round(d['ms_per_step'],3),'ondev ms',round(d['ms_per_step_results_on_device'],3),d['value_results_on_device_kernels_ms_per_step'],d['roofline']['all_kernels_ms_per_step'])
"
done
