mkdir -p gpurun_out/r2y
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 --no-north-star > gpurun_out/r2y/bench$N.json 2> gpurun_out/r2y/bench$N.err
timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/r2y/bench1.json 2> gpurun_out/r2y/bench1.err
python - <<PY
import json
for f in ('bench1','bench$N'):
    try:
        txt=[l for l in open(f'gpurun_out/r2y/{f}.json') if l.startswith('{')][-1]
        d=json.loads(txt)
        print(f, d.get('n_gpus'), 'steps', d.get('steps'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'graphs', d.get('config',{}).get('graphs'), (d.get('parity_check') or {}).get('ok'), d.get('config',{}).get('ms_first_30_steps'), 'wall', d.get('config',{}).get('wall_ms_per_step'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -2 gpurun_out/r2y/bench$N.err
