set -x
mkdir -p gpurun_out/r2r
nvidia-smi topo -m > gpurun_out/r2r/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 > gpurun_out/r2r/bench8.json 2> gpurun_out/r2r/bench8.err
tail -3 gpurun_out/r2r/bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 > gpurun_out/r2r/bench4.json 2> gpurun_out/r2r/bench4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --no-north-star > gpurun_out/r2r/bench2.json 2> gpurun_out/r2r/bench2.err
timeout 300 python bench.py --no-extras > gpurun_out/r2r/bench1.json 2> gpurun_out/r2r/bench1.err
python - <<'PY'
import json
for f in ('bench1','bench2','bench4','bench8'):
    try:
        txt=[l for l in open(f'gpurun_out/r2r/{f}.json') if l.startswith('{')][-1]
        d=json.loads(txt)
        print(f, d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'graphs', d.get('config',{}).get('graphs'), d.get('config',{}).get('ms_first_30_steps'), d.get('config',{}).get('ms_last_30_steps'))
        for k in ('north_star','strong_scaling','parity_check'):
            if k in d: print('   ', k, json.dumps(d[k])[:500])
    except Exception as e: print(f, 'ERR', e)
PY
