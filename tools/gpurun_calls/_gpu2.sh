set -x
mkdir -p gpurun_out/r2b
timeout 1400 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2b/pytest_gpu.txt 2>&1
tail -30 gpurun_out/r2b/pytest_gpu.txt
