set -x
mkdir -p gpurun_out/r2j
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2j/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2j/pytest_gpu.txt
python tools/profile_box.py --cross 512 --depth 64 --steps 6 > gpurun_out/r2j/prof_16m.json 2>&1
python tools/profile_box.py --cross 100 --depth 100 --steps 30 > gpurun_out/r2j/prof_1m.json 2>&1
cat gpurun_out/r2j/prof_16m.json gpurun_out/r2j/prof_1m.json
python tools/bench_kernels.py > gpurun_out/r2j/kernels.jsonl 2> gpurun_out/r2j/kernels.err
cat gpurun_out/r2j/kernels.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'agent_function_wrapper|k_scan_scatter|k_radix_onesweep|k_bin_scatter_staged|k_sort_keys_hist|k_group_tile|k_gather' -s 26 -c 13 -o gpurun_out/r2j/step16m python tools/run_circles.py --n 16777216 --steps 4 --iter-mode -1 > gpurun_out/r2j/ncu16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k agent_function_wrapper -s 20 -c 2 -o gpurun_out/r2j/funcs_1m_step10 python tools/run_circles.py --steps 12 --iter-mode -1 > gpurun_out/r2j/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j/launches_1m.csv python tools/run_circles.py --steps 12 --graphs 0 --iter-mode -1 > gpurun_out/r2j/ncul.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j/launches_16m.csv python tools/run_circles.py --n 16777216 --steps 4 --graphs 0 --iter-mode -1 > gpurun_out/r2j/ncul16.log 2>&1
ls -la gpurun_out/r2j
