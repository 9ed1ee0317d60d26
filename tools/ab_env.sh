#!/bin/bash
# tools/ab_env.sh VAR v1 v2 ...: device-resident bench once per value of an environment variable
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 150 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$var=$v', round(d['value']), round(d['ms_per_step'],3), {k:round(x['ms_per_step'],3) for k,x in d['kernels'].items()})"
done
