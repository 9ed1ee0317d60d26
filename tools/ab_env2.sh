#!/bin/bash
# tools/ab_env2.sh VAR v1 v2 ...: the device-resident bench with environment variable VAR set to each value in turn (twice)
V=$1; shift
for rep in 1 2; do for x in "$@"; do
  env $V=$x timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$V=$x', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done; done
