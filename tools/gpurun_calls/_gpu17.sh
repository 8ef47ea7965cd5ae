set -x
mkdir -p gpurun_out/r2q
SLAB_STEPS=12 timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2q/pytest_slab.txt 2>&1
tail -3 gpurun_out/r2q/pytest_slab.txt | cut -c1-600
for d in 64 100; do
c=512; [ $d = 100 ] && c=100
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/profile_slab.py --cross $c --depth $d > gpurun_out/r2q/prof2_$d.jsonl 2> gpurun_out/r2q/prof2_$d.err
grep '^{' gpurun_out/r2q/prof2_$d.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/profile_slab.py --cross $c --depth $d --profile 0 --steps 30 > gpurun_out/r2q/graph2_$d.jsonl 2> gpurun_out/r2q/graph2_$d.err
grep '^{' gpurun_out/r2q/graph2_$d.jsonl
python tools/profile_box.py --cross $c --depth $d --steps 8 | tee gpurun_out/r2q/prof1_$d.json
done
timeout 600 python -m pytest tests/test_sim_gpu.py tests/test_semantics_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
