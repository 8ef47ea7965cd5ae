mkdir -p gpurun_out/r2ab
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2ab/pytest_gpu.txt 2>&1
tail -4 gpurun_out/r2ab/pytest_gpu.txt | cut -c1-400
python tools/profile_box.py --cross 512 --depth 64 --steps 6 | tee gpurun_out/r2ab/prof_16m.json
python tools/profile_box.py --cross 100 --depth 100 --steps 30 | tee gpurun_out/r2ab/prof_1m.json
timeout 300 python tests/bench_configs.py --only boids2d_16m --no-ref | cut -c1-700 | tee gpurun_out/r2ab/boids2d.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras | cut -c1-300
