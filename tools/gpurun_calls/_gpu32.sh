mkdir -p gpurun_out/r2final3
timeout 1200 python bench.py > gpurun_out/r2final3/bench1.json 2> gpurun_out/r2final3/bench1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final3/bench1_k20.json 2> gpurun_out/r2final3/bench1_k20.err
timeout 300 python tests/bench_configs.py --no-ref > gpurun_out/r2final3/configs.jsonl 2> gpurun_out/r2final3/configs.err
python - <<'PY'
import json
for f in ('bench1','bench1_k20'):
    d=json.load(open(f'gpurun_out/r2final3/{f}.json'))
    print(f, 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'roofline', round(d['roofline']['frac'],3), d['roofline']['us_per_launch'], {k: round(v['frac'],3) for k,v in d['rooflines'].items()}, d.get('north_star',{}).get('ms_per_step'), d.get('strong_scaling',{}).get('ms_per_step'), d.get('reference_cuda',{}).get('ms_per_step'))
for l in open('gpurun_out/r2final3/configs.jsonl'):
    d=json.loads(l); print(d['config'], round(d['ms_per_step'],3), {k: round(v) for k,v in d['phases_us'].items()})
PY
