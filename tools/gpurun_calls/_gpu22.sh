set -x
mkdir -p gpurun_out/r2v
SLAB_STEPS=12 timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2v/pytest_slab.txt 2>&1
tail -6 gpurun_out/r2v/pytest_slab.txt | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --no-north-star --steps 20 --warmup 5 > gpurun_out/r2v/bench2.json 2> gpurun_out/r2v/bench2.err
timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/r2v/bench1.json 2> gpurun_out/r2v/bench1.err
timeout 600 python -m pytest tests/test_sim_gpu.py tests/test_semantics_gpu.py tests/test_ref_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
python - <<'PY'
import json
for f in ('bench1','bench2'):
    try:
        txt=[l for l in open(f'gpurun_out/r2v/{f}.json') if l.startswith('{')][-1]
        d=json.loads(txt)
        print(f, d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'graphs', d.get('config',{}).get('graphs'), (d.get('parity_check') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2v/bench2.err
