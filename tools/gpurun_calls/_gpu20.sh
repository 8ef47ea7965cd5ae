set -x
mkdir -p gpurun_out/r2t
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2t/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2t/pytest_gpu.txt | cut -c1-400
for v in 1 0; do
python tools/run_circles.py --steps 330 --graphs 1 --iter-mode -1 --times --extra validation=$v > gpurun_out/r2t/times_val$v.txt 2>&1
echo validation=$v; cat gpurun_out/r2t/times_val$v.txt
done
timeout 900 python tests/bench_configs.py --only circles3d_16m --no-ref > gpurun_out/r2t/configs.jsonl 2> gpurun_out/r2t/configs.err
cut -c1-900 gpurun_out/r2t/configs.jsonl
python tools/profile_box.py --cross 512 --depth 512 --steps 4 | tee gpurun_out/r2t/prof_128m.json
