#!/bin/bash
# tools/build_variant.sh NAME [-DMACRO=..]...: builds dmmt_jpeg_encoder_b200/lib/libdmmt_NAME.so with extra macros for
# k1_transform.cu / k2_entropy.cu (A/B experiments; select at run time with DMMT_CUDA_LIB, see tools/ab_bench.sh)
set -e
name=$1; shift
cd "$(dirname "$0")/../dmmt_jpeg_encoder_b200/csrc"
mkdir -p /tmp/v_$name
for f in k1_transform k2_entropy; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC "$@" -Xptxas -v -c $f.cu -o /tmp/v_$name/$f.o 2>&1 | grep -A2 "p420ILi1ELb1\|k3_pack" | grep "Used\|spill" || true
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../lib/libdmmt_$name.so /tmp/v_$name/k1_transform.o /tmp/v_$name/k2_entropy.o ../lib/obj/dmmt_api.o ../lib/obj/dmmt_batch.o ../lib/obj/dmmt_shard.o
echo built libdmmt_$name.so
