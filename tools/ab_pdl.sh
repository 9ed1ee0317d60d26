for L in libdmmt_cuda.so libdmmt_notrig.so; do for P in 0 1; do
  DMMT_CUDA_LIB=$PWD/dmmt_jpeg_encoder_b200/lib/$L DMMT_PDL=$P timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$L PDL=$P', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, round(d['extra']['config3']['device_us'],1))"
done; done
