set -x
mkdir -p gpurun_out/r2p
for d in 64 50; do
c=512; [ $d = 50 ] && c=100
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/profile_slab.py --cross $c --depth $d > gpurun_out/r2p/prof2_$d.jsonl 2> gpurun_out/r2p/prof2_$d.err
grep '^{' gpurun_out/r2p/prof2_$d.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/profile_slab.py --cross $c --depth $d --profile 0 > gpurun_out/r2p/graph2_$d.jsonl 2> gpurun_out/r2p/graph2_$d.err
grep '^{' gpurun_out/r2p/graph2_$d.jsonl
python tools/profile_box.py --cross $c --depth $d --steps 8 | tee gpurun_out/r2p/prof1_$d.json
done
tail -3 gpurun_out/r2p/*.err
