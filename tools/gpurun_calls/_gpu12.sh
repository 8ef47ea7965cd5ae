set -x
mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2l/pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2l/pytest_gpu.txt | cut -c1-300; grep -a "dropin\]" gpurun_out/r2l/pytest_gpu.txt
FGB_COMPACT_BULK=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_semantics_gpu.py tests/test_ref_parity_gpu.py -m gpu -q --timeout 900 > gpurun_out/r2l/pytest_bulk.txt 2>&1
tail -4 gpurun_out/r2l/pytest_bulk.txt | cut -c1-300
for b in 0 1; do
  FGB_COMPACT_BULK=$b python tools/bench_kernels.py > gpurun_out/r2l/kernels_bulk$b.jsonl 2> gpurun_out/r2l/kernels_bulk$b.err
  echo bulk=$b; grep compact gpurun_out/r2l/kernels_bulk$b.jsonl | cut -c1-330
done
for b in 0 1; do
FGB_COMPACT_BULK=$b timeout 600 ncu --set full --clock-control none --import-source on -k k_compact -s 4 -c 1 -o gpurun_out/r2l/compact_bulk$b python tools/bench_kernels.py --sizes 16777216 --reps 3 > gpurun_out/r2l/ncu_c$b.log 2>&1
done
ls -la gpurun_out/r2l
