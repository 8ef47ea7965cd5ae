mkdir -p gpurun_out/r2final2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final2/bench1_k20.json 2> gpurun_out/r2final2/bench1_k20.err
tail -2 gpurun_out/r2final2/bench1_k20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2final2/ref1_k20.json 2> gpurun_out/r2final2/ref1_k20.err
python - <<'PY'
import json
for f in ('bench1_k20','ref1_k20'):
    d=json.load(open(f'gpurun_out/r2final2/{f}.json'))
    print(f, 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), sorted(d.keys()))
    print('   config keys', sorted(d.get('config',{}).keys()))
PY
timeout 900 python -m pytest tests/test_sim_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 600 2>&1 | tail -2
