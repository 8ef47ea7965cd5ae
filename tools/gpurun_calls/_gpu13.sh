set -x
mkdir -p gpurun_out/r2m
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_semantics_gpu.py tests/test_ref_parity_gpu.py tests/test_sim_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/r2m/pytest_gpu.txt 2>&1
tail -4 gpurun_out/r2m/pytest_gpu.txt | cut -c1-300
for v in main c4; do
  if [ $v = main ]; then unset FGB_KERNELS_LIB; else export FGB_KERNELS_LIB=$PWD/flamegpu2_b200/lib/ab/libflamegpu2_b200_$v.so; fi
  python tools/bench_kernels.py > gpurun_out/r2m/kernels_$v.jsonl 2> gpurun_out/r2m/kernels_$v.err
  echo $v; grep compact gpurun_out/r2m/kernels_$v.jsonl | cut -c1-330
done
unset FGB_KERNELS_LIB
# ncu: build kernels inside the step (north-star box and 1 M cube), all step kernels once
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scan_scatter|k_bin_scatter_staged|k_radix_onesweep|k_sort_keys_hist|k_group_tile|k_gather|agent_function_wrapper' -s 24 -c 12 -o gpurun_out/r2m/step_ns16m python tools/run_circles.py --cross 512 --depth 64 --steps 4 --iter-mode -1 > gpurun_out/r2m/ncu16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scan_scatter|k_bin_scatter_staged|k_radix_onesweep|k_sort_keys_hist|k_group_tile|k_gather|agent_function_wrapper' -s 108 -c 12 -o gpurun_out/r2m/step_1m python tools/run_circles.py --steps 12 --iter-mode -1 > gpurun_out/r2m/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m/launches_ns16m.csv python tools/run_circles.py --cross 512 --depth 64 --steps 4 --graphs 0 --iter-mode -1 > gpurun_out/r2m/ncul16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m/launches_1m.csv python tools/run_circles.py --steps 12 --graphs 0 --iter-mode -1 > gpurun_out/r2m/ncul1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k agent_function_wrapper -s 601 -c 1 -o gpurun_out/r2m/move_step300 python tools/run_circles.py --steps 302 --iter-mode -1 > gpurun_out/r2m/ncu300.log 2>&1
ls -la gpurun_out/r2m
