for cfg in "128 3" "128 4" "64 3" "64 4" "64 6" "32 6" "256 2" "128 2"; do set -- $cfg
python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 10 --sub-batch $1 --depth $2 --exact-sub-batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sub $1 depth $2', round(d['value']), round(d['ms_per_step'],3))"
done
