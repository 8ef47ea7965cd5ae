mkdir -p gpurun_out/r2final
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2final/pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2final/pytest_gpu.txt | cut -c1-300; grep -a "dropin\]" gpurun_out/r2final/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2final/bench1.json 2> gpurun_out/r2final/bench1.err
tail -3 gpurun_out/r2final/bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2final/ref1.json 2> gpurun_out/r2final/ref1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2final/bench1_k20.json 2> gpurun_out/r2final/bench1_k20.err
python - <<'PY'
import json
for f in ('bench1','ref1','bench1_k20'):
    try:
        d=json.load(open(f'gpurun_out/r2final/{f}.json'))
        print(f, d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'graphs', d.get('config',{}).get('graphs'), d.get('config',{}).get('ms_first_30_steps'), d.get('config',{}).get('ms_last_30_steps'))
        for k in ('roofline','north_star','strong_scaling','phases_us','phases_us_16m','modes','reference_cuda','move'):
            if k in d: print('   ', k, json.dumps(d[k])[:600])
        if 'rooflines' in d:
            for k,v in d['rooflines'].items(): print('   rl', k, round(v.get('frac',0),3), round(v.get('us_per_launch',0),1))
    except Exception as e: print(f, 'ERR', e)
PY
