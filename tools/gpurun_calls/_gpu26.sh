mkdir -p gpurun_out/r2z
N=4
for mode in 0 1 0 1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus $N --steps 20 --warmup 5 --no-extras --py-loop $mode > gpurun_out/r2z/bench${N}_py$mode.json 2> gpurun_out/r2z/bench${N}_py$mode.err
python - <<PY
import json
txt=[l for l in open('gpurun_out/r2z/bench${N}_py$mode.json') if l.startswith('{')][-1]
d=json.loads(txt)
print('py_loop=$mode', d.get('n_gpus'), 'ms/step', round(d.get('ms_per_step'),4), 'rank0 per-step', round(d.get('config',{}).get('ms_first_30_steps'),4), 'wall', round(d.get('config',{}).get('wall_ms_per_step'),4), d['clocks'])
PY
done
