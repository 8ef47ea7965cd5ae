mkdir -p gpurun_out/r2ac
python tools/bench_kernels.py --reps 10 > gpurun_out/r2ac/kernels.jsonl 2> gpurun_out/r2ac/kernels.err
python - <<'PY'
import json
for l in open('gpurun_out/r2ac/kernels.jsonl'):
    d=json.loads(l)
    if d['op']=='build_index': print(d['n'], d['input'], 'stable' if d['stable'] else '', 'expect' if d.get('expect_grouped') else '', round(d['us_median'],1), round(d['frac_of_measured_peak'],3))
PY
python tools/bench_kernels.py --bucket --reps 5 | cut -c1-260
