set -x
mkdir -p gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2n/pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2n/pytest_gpu.txt | cut -c1-300; grep -a "dropin\]" gpurun_out/r2n/pytest_gpu.txt
python tools/bench_kernels.py > gpurun_out/r2n/kernels.jsonl 2> gpurun_out/r2n/kernels.err
grep -v '"stable": true' gpurun_out/r2n/kernels.jsonl | cut -c1-330
timeout 1200 python bench.py > gpurun_out/r2n/bench1.json 2> gpurun_out/r2n/bench1.err
tail -3 gpurun_out/r2n/bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2n/ref1.json 2> gpurun_out/r2n/ref1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scan_scatter|k_bin_scatter_staged|k_radix_onesweep|k_sort_keys_hist|k_group_tile|k_gather|agent_function_wrapper' -s 60 -c 12 -o gpurun_out/r2n/step_1m python tools/run_circles.py --steps 12 --iter-mode -1 > gpurun_out/r2n/ncu1.log 2>&1
python - <<'PY'
import json
for f in ('bench1','ref1'):
    try:
        d=json.load(open(f'gpurun_out/r2n/{f}.json'))
        print(f, d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'graphs', d.get('config',{}).get('graphs'), d.get('config',{}).get('ms_first_30_steps'), d.get('config',{}).get('ms_last_30_steps'))
        for k in ('roofline','rooflines','north_star','strong_scaling','phases_us','phases_us_16m','modes','reference_cuda'):
            if k in d: print('   ', k, json.dumps(d[k])[:900])
    except Exception as e: print(f, 'ERR', e)
PY
ls -la gpurun_out/r2n
