set -x
mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2g/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2g/pytest_gpu.txt
python tools/profile_box.py --cross 512 --depth 64 > gpurun_out/r2g/prof_16m.json 2>&1
python tools/profile_box.py --cross 512 --depth 512 --steps 4 > gpurun_out/r2g/prof_128m.json 2>&1
python tools/profile_box.py --cross 512 --depth 512 --steps 4 --iter-mode 0 > gpurun_out/r2g/prof_128m_mode0.json 2>&1
cat gpurun_out/r2g/prof_*.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2g/bench2.json 2> gpurun_out/r2g/bench2.err
tail -5 gpurun_out/r2g/bench2.err
grep -v "^NCCL" gpurun_out/r2g/bench2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=2', d['ms_per_step'], d['value'], d['parity_check'])
print(d.get('north_star')); print(d.get('strong_scaling'))
"
timeout 300 python bench.py --no-extras > gpurun_out/r2g/bench1.json 2> gpurun_out/r2g/bench1.err
python -c "
import json
d=json.load(open('gpurun_out/r2g/bench1.json')); print('N=1', d['ms_per_step'], d['value'])
"
