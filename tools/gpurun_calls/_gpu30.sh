mkdir -p gpurun_out/r2aa
for b in 64 128 192 256; do
timeout 200 python bench.py --steps 100 --warmup 5 --no-extras --block $b > gpurun_out/r2aa/bench_b$b.json 2> gpurun_out/r2aa/bench_b$b.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2aa/bench_b$b.json'))
print('block $b', 'ms/step', round(d['ms_per_step'],4), 'first30', round(d['config']['ms_first_30_steps'],4), 'last30', round(d['config']['ms_last_30_steps'],4))
PY
done
