#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -15 $O/pytest.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | tail -4; tail -3 $O/bench_default.err
python - <<'P'
import json
for line in open('gpurun_out/r2e/bench_default.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('headline', round(d['ms_per_step'],4), '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['gpu_launches'], d['roofline'] and {k:d['roofline'][k] for k in ('kernel','bound','frac','achieved')}, d['cpu_baseline'])
        x=d.get('parity_mode_bf16x3'); print('x3', x and (round(x['ms_per_step'],4), '%.4g'%x['value'], x['roofline'] and {k:x['roofline'][k] for k in ('kernel','bound','frac','achieved')}))
        print('predict', d.get('predict'))
        w=d.get('roofline_wind_map_e16'); print('wind', w and (round(w['ms_per_step'],3), w['roofline'], w['clocks'], {k:round(v['ms_per_step'],3) for k,v in w['kernels'].items()}))
P
for a in "--precision bf16 --workload air_quality_mle_zinb_e8 --steps 10 --warmup 4" "--precision bf16x3 --workload air_quality_mle_zinb_e8 --steps 4 --warmup 2"; do
timeout 300 python bench.py $a --no-cpu-baseline 2> $O/bench.err | tee -a $O/bench.jsonl | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$a', round(d['ms_per_step'],4), '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['gpu_launches'], d['timing']['blocks'], {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
tail -3 $O/bench.err
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 | cut -c1-400
