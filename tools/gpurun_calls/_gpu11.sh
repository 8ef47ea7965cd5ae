set -x
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x -s > gpurun_out/r2k/pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2k/pytest_gpu.txt; grep dropin gpurun_out/r2k/pytest_gpu.txt
for v in main v5 v6; do
  if [ $v = main ]; then unset FGB_KERNELS_LIB; else export FGB_KERNELS_LIB=$PWD/flamegpu2_b200/lib/ab/libflamegpu2_b200_$v.so; fi
  python tools/bench_kernels.py --sizes 16777216 > gpurun_out/r2k/kernels_$v.jsonl 2> gpurun_out/r2k/kernels_$v.err
  echo $v; grep -v '"stable": true' gpurun_out/r2k/kernels_$v.jsonl | grep -v random | cut -c1-330
done
unset FGB_KERNELS_LIB
python tools/profile_box.py --cross 512 --depth 64 --steps 6 > gpurun_out/r2k/prof_16m.json 2>&1
python tools/profile_box.py --cross 100 --depth 100 --steps 30 > gpurun_out/r2k/prof_1m.json 2>&1
cat gpurun_out/r2k/prof_16m.json gpurun_out/r2k/prof_1m.json
ls -la gpurun_out/r2k
