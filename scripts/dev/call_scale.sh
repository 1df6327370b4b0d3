#!/bin/bash
# usage: call_scale.sh N  -- the driver's scaling command at N ranks
N=$1; O=gpurun_out/scale; mkdir -p $O
if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py"; fi
( time timeout 600 $CMD --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err ) 2>&1 | tail -3
tail -5 $O/bench_n$N.err
python - "$O/bench_n$N.json" <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['timing'], d['clocks'])
        x=d.get('parity_mode_bf16x3'); print('x3', x and (round(x['ms_per_step'],4), '%.4g'%x['value']))
        print('predict', d.get('predict'))
        w=d.get('roofline_wind_map_e16'); print('wind', w and (round(w['ms_per_step'],3), '%.4g'%w['value'], w['clocks']))
P
