#!/bin/bash
# like tools/ab_bench.sh, plus the single-4K-frame chain (extra.config3): device us and per-kernel us
for L in "$@"; do
  DMMT_CUDA_LIB=$PWD/dmmt_jpeg_encoder_b200/lib/$L timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$L', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, round(d['extra']['config3']['device_us'],1), d['extra']['config3']['kernel_us'])"
done
