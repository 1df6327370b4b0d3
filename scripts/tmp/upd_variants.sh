for v in 0 1 2 4 7 8; do
  for wl in chickenpox_map_e8 wind_map_e16; do
    BNF_UPD_VARIANT=$v timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('variant $v $wl', d['ms_per_step'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('map_update','encode','head_fused')})
"
  done
done
