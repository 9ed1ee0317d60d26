for cfg in "256 2" "342 3" "512 2" "205 2" "171 3" "256 3" "128 2"; do set -- $cfg
python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 10 --images 1024 --sub-batch $1 --depth $2 --exact-sub-batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('images 1024 sub $1 depth $2', round(d['value']), round(d['ms_per_step'],4))"
done
