#!/bin/bash
# A/B of library builds with the same ABI on one GPU box: tools/ab_bench.sh libdmmt_cuda.so libdmmt_x.so ...
# (each library is selected with DMMT_CUDA_LIB; prints device-resident MPixel/s, ms per step and per-kernel ms)
for L in "$@"; do
  DMMT_CUDA_LIB=$PWD/dmmt_jpeg_encoder_b200/lib/$L timeout 150 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$L', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done
