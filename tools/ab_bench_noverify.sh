for L in "$@"; do
  DMMT_CUDA_LIB=$PWD/dmmt_jpeg_encoder_b200/lib/$L timeout 150 python bench.py --no-cpu-baseline --no-e2e --no-extra --no-verify --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$L', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done
