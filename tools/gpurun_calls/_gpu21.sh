set -x
mkdir -p gpurun_out/r2u
for m in 0 1; do
timeout 600 python tests/bench_configs.py --only boids2d_16m --no-ref --iter-mode $m > gpurun_out/r2u/boids2d_mode$m.jsonl 2>&1
cut -c1-700 gpurun_out/r2u/boids2d_mode$m.jsonl
done
