set -x
mkdir -p gpurun_out/r2d
timeout 1400 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2d/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2d/pytest_gpu.txt
python bench.py --no-extras --steps 100 --iter-mode -1 > gpurun_out/r2d/bench_1m.json 2> gpurun_out/r2d/bench_1m.err
python bench.py --no-extras --steps 30 --iter-mode -1 --agents-per-gpu 16777216 > gpurun_out/r2d/bench_16m.json 2> gpurun_out/r2d/bench_16m.err
python - <<'PY'
import json
for f in ("gpurun_out/r2d/bench_1m.json","gpurun_out/r2d/bench_16m.json"):
    try:
        d=json.load(open(f)); print(f, d["ms_per_step"], d["phases_us"])
    except Exception as e: print(f, e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d/launches_1m.csv python tools/run_circles.py --steps 12 --graphs 0 --iter-mode -1 > gpurun_out/r2d/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d/launches_16m.csv python tools/run_circles.py --n 16777216 --steps 6 --graphs 0 --iter-mode -1 > gpurun_out/r2d/ncu16.log 2>&1
