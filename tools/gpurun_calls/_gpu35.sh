mkdir -p gpurun_out/r2ad
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2ad/pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2ad/pytest_gpu.txt | cut -c1-400
