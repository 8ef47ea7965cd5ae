mkdir -p gpurun_out/r2final4
SLAB_STEPS=12 timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2final4/pytest_slab.txt 2>&1
tail -3 gpurun_out/r2final4/pytest_slab.txt | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2final4/bench2_k20.json 2> gpurun_out/r2final4/bench2_k20.err
python - <<PY
import json
txt=[l for l in open('gpurun_out/r2final4/bench2_k20.json') if l.startswith('{')][-1]
d=json.loads(txt)
print(d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'region', d['config'].get('region_ms_per_step'), 'value', d.get('value'), (d.get('parity_check') or {}).get('ok'), d.get('north_star',{}).get('ms_per_step'), d.get('strong_scaling',{}).get('ms_per_step'))
PY
tail -2 gpurun_out/r2final4/bench2_k20.err
