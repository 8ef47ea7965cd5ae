set -x
mkdir -p gpurun_out/r2s
for v in 1 0; do
python tools/run_circles.py --steps 330 --graphs 1 --iter-mode -1 --times --extra validation=$v > gpurun_out/r2s/times_val$v.txt 2>&1
echo validation=$v; cat gpurun_out/r2s/times_val$v.txt
done
timeout 1500 python tests/bench_configs.py > gpurun_out/r2s/configs.jsonl 2> gpurun_out/r2s/configs.err
cut -c1-900 gpurun_out/r2s/configs.jsonl
tail -3 gpurun_out/r2s/configs.err
