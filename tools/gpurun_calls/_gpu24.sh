mkdir -p gpurun_out/r2x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/profile_slab.py --cross 100 --depth 100 --profile 0 --steps 44 > gpurun_out/r2x/graph8.jsonl 2> gpurun_out/r2x/graph8.err
grep -h '^{' gpurun_out/r2x/graph8.jsonl | sed 's/}{/}\n{/g'
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/profile_slab.py --cross 100 --depth 100 --profile 1 --steps 20 > gpurun_out/r2x/prof8.jsonl 2> gpurun_out/r2x/prof8.err
grep -h '^{' gpurun_out/r2x/prof8.jsonl | sed 's/}{/}\n{/g' | cut -c1-600
