set -x
mkdir -p gpurun_out/r2e
nvidia-smi topo -m > gpurun_out/r2e/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/slab_parity_check.py > gpurun_out/r2e/parity2.txt 2>&1
tail -12 gpurun_out/r2e/parity2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --iter-mode -1 > gpurun_out/r2e/bench2.json 2> gpurun_out/r2e/bench2.err
tail -3 gpurun_out/r2e/bench2.err; cat gpurun_out/r2e/bench2.json
timeout 300 python bench.py --no-extras --steps 200 --iter-mode -1 > gpurun_out/r2e/bench1.json 2> gpurun_out/r2e/bench1.err
python -c "
import json
for f in ('gpurun_out/r2e/bench1.json','gpurun_out/r2e/bench2.json'):
    try:
        d=json.load(open(f)); print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['config'].get('graphs'))
    except Exception as e: print(f, e)
"
timeout 600 python -m pytest tests/test_sim_gpu.py tests/test_kernels_gpu.py -q -x --timeout 300 2>&1 | tail -5
