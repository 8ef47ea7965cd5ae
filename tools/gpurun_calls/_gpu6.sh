set -x
mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2f/pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2f/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2f/bench1.json 2> gpurun_out/r2f/bench1.err
tail -3 gpurun_out/r2f/bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2f/ref1.json 2> gpurun_out/r2f/ref1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2f/bench2.json 2> gpurun_out/r2f/bench2.err
tail -5 gpurun_out/r2f/bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 > gpurun_out/r2f/ref2.json 2> gpurun_out/r2f/ref2.err
python - <<'PY'
import json
for f in ('bench1','ref1','bench2','ref2'):
    try:
        d=json.load(open(f'gpurun_out/r2f/{f}.json'))
        print(f, d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'graphs', d.get('config',{}).get('graphs'))
        for k in ('roofline','rooflines','north_star','strong_scaling','phases_us','phases_us_16m','modes','reference_cuda','parity_check'):
            if k in d: print('   ', k, json.dumps(d[k])[:700])
    except Exception as e: print(f, 'ERR', e)
PY
