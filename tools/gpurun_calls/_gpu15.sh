set -x
mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2o/pytest_slab.txt 2>&1
tail -3 gpurun_out/r2o/pytest_slab.txt | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2o/bench2.json 2> gpurun_out/r2o/bench2.err
tail -5 gpurun_out/r2o/bench2.err
timeout 600 python bench.py --no-extras > gpurun_out/r2o/bench1.json 2> gpurun_out/r2o/bench1.err
python - <<'PY'
import json
for f in ('bench1','bench2'):
    try:
        txt=[l for l in open(f'gpurun_out/r2o/{f}.json') if l.startswith('{')][-1]
        d=json.loads(txt)
        print(f, d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'graphs', d.get('config',{}).get('graphs'), d.get('config',{}).get('ms_first_30_steps'), d.get('config',{}).get('ms_last_30_steps'))
        for k in ('north_star','strong_scaling','parity_check'):
            if k in d: print('   ', k, json.dumps(d[k])[:700])
    except Exception as e: print(f, 'ERR', e)
PY
