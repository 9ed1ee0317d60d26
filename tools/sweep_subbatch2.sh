for cfg in "1024 256 3" "1024 512 2" "1024 342 3" "1024 256 2" "128 128 1" "128 64 2" "128 43 3" "128 64 3" "256 128 2" "256 86 3" "512 256 2" "512 171 3"; do set -- $cfg
python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 10 --images $1 --sub-batch $2 --depth $3 --exact-sub-batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('images $1 sub $2 depth $3', round(d['value']), round(d['ms_per_step'],3))"
done
