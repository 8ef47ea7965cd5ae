set -x
mkdir -p gpurun_out/r2a
for m in 0 1; do
timeout 900 ncu --set full --clock-control none --import-source on -k agent_function_wrapper -s 601 -c 1 -o gpurun_out/r2a/move_mode${m}_step300 python tools/run_circles.py --steps 301 --iter-mode $m > gpurun_out/r2a/ncu_mode$m.log 2>&1
done
ls -la gpurun_out/r2a
