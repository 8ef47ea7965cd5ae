set -x
mkdir -p gpurun_out/r2h
python tools/profile_box.py --cross 512 --depth 512 --steps 4 > gpurun_out/r2h/prof_128m.json 2>&1
python tools/profile_box.py --cross 512 --depth 512 --steps 4 --iter-mode 0 > gpurun_out/r2h/prof_128m_mode0.json 2>&1
python tools/profile_box.py --cross 512 --depth 64 --steps 6 > gpurun_out/r2h/prof_16m.json 2>&1
python tools/profile_box.py --cross 512 --depth 64 --steps 6 --extra tile_order=0 > gpurun_out/r2h/prof_16m_global.json 2>&1
python tools/profile_box.py --cross 100 --depth 100 --steps 30 --extra tile_order=0 > gpurun_out/r2h/prof_1m_global.json 2>&1
python tools/profile_box.py --cross 100 --depth 100 --steps 30 > gpurun_out/r2h/prof_1m.json 2>&1
cat gpurun_out/r2h/prof_*.json
# ncu: the step's kernels at 16.8 M (step 3 of an eager run) and `move` at 1 M, step 10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_scatter_direct|k_exclusive_scan|k_radix_onesweep|k_gather|k_sort_keys_hist|k_group_tile' -s 18 -c 7 -o gpurun_out/r2h/step16m python tools/run_circles.py --n 16777216 --steps 4 --iter-mode -1 > gpurun_out/r2h/ncu16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k agent_function_wrapper -s 20 -c 2 -o gpurun_out/r2h/funcs_1m_step10 python tools/run_circles.py --steps 12 --iter-mode -1 > gpurun_out/r2h/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h/launches_1m.csv python tools/run_circles.py --steps 12 --graphs 0 --iter-mode -1 > gpurun_out/r2h/ncul.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -4
ls -la gpurun_out/r2h
