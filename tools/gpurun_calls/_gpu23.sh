set -x
mkdir -p gpurun_out/r2w
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-north-star > gpurun_out/r2w/bench8.json 2> gpurun_out/r2w/bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 5 --no-north-star > gpurun_out/r2w/bench4.json 2> gpurun_out/r2w/bench4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 300 --warmup 10 --no-extras > gpurun_out/r2w/bench8_300.json 2> gpurun_out/r2w/bench8_300.err
timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/r2w/bench1.json 2> gpurun_out/r2w/bench1.err
python - <<'PY'
import json
for f in ('bench1','bench4','bench8','bench8_300'):
    try:
        txt=[l for l in open(f'gpurun_out/r2w/{f}.json') if l.startswith('{')][-1]
        d=json.loads(txt)
        print(f, d.get('n_gpus'), 'steps', d.get('steps'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'graphs', d.get('config',{}).get('graphs'), (d.get('parity_check') or {}).get('ok'), d.get('config',{}).get('ms_first_30_steps'), d.get('config',{}).get('ms_last_30_steps'))
    except Exception as e: print(f, 'ERR', e)
PY
