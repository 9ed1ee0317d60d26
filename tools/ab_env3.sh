#!/bin/bash
# tools/ab_env3.sh IMAGES VAR v1 v2 ...: like ab_env2.sh with --images IMAGES (e.g. 128 = one rank's share at 8 GPUs)
N=$1; V=$2; shift 2
for rep in 1 2; do for x in "$@"; do
  env $V=$x timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 20 --images $N 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('images $N $V=$x', round(d['value']), round(d['ms_per_step'],4))"
done; done
