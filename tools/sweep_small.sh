for cfg in "128 1" "64 2" "64 3" "43 3" "32 2" "32 4" "22 3" "16 4"; do set -- $cfg
python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 30 --images 128 --sub-batch $1 --depth $2 --exact-sub-batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('images 128 sub $1 depth $2', round(d['value']), round(d['ms_per_step'],4))"
done
