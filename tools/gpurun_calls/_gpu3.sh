set -x
mkdir -p gpurun_out/r2c
timeout 1400 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2c/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2c/pytest_gpu.txt
python bench.py --no-extras --steps 100 > gpurun_out/r2c/bench_1m.json 2> gpurun_out/r2c/bench_1m.err
python bench.py --no-extras --steps 30 --agents-per-gpu 16777216 > gpurun_out/r2c/bench_16m.json 2> gpurun_out/r2c/bench_16m.err
cat gpurun_out/r2c/bench_1m.json gpurun_out/r2c/bench_16m.json
