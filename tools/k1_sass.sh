#!/bin/bash
# tools/k1_sass.sh [-DMACRO=..]...: compiles k1_transform.cu with extra macros and prints, for the fused u8 4:2:0 kernel
# (k1_transform_p420<u8, fused, aligned>), registers / spills and the SASS opcode histogram; the listing goes to /tmp/k1.sass
cd "$(dirname "$0")/../dmmt_jpeg_encoder_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false "$@" -Xptxas -v -c k1_transform.cu -o /tmp/k1_variant.o 2>&1 | grep -A2 "p420ILi1ELb1ELb1" | grep -E "Used|spill"
cuobjdump -sass /tmp/k1_variant.o | awk '/Function : .*p420ILi1ELb1ELb1/{f=1;next} f&&/Function :/{f=0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\* 0x[0-9a-f]* \*\///' > /tmp/k1.sass
echo "instructions: $(wc -l < /tmp/k1.sass)"
awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); c[op]++} END{for(o in c) printf "%s:%d ", o, c[o]; print ""}' /tmp/k1.sass
