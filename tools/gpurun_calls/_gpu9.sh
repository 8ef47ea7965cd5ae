set -x
mkdir -p gpurun_out/r2i
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2i/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2i/pytest_gpu.txt
for v in base new q10 q12; do
  if [ $v = new ]; then unset FGB_MODELS_LIB; else export FGB_MODELS_LIB=$PWD/flamegpu2_b200/lib/ab/libfgb_models_$v.so; fi
  python tools/run_circles.py --steps 330 --graphs 1 --iter-mode -1 --times > gpurun_out/r2i/times_$v.txt 2>&1
  echo $v; cat gpurun_out/r2i/times_$v.txt
done
for v in base new q10; do
  if [ $v = new ]; then unset FGB_MODELS_LIB; else export FGB_MODELS_LIB=$PWD/flamegpu2_b200/lib/ab/libfgb_models_$v.so; fi
  python tools/profile_box.py --cross 512 --depth 64 --steps 6 > gpurun_out/r2i/prof_16m_$v.json 2>&1
  python tools/profile_box.py --cross 100 --depth 100 --steps 30 > gpurun_out/r2i/prof_1m_$v.json 2>&1
  cat gpurun_out/r2i/prof_16m_$v.json gpurun_out/r2i/prof_1m_$v.json
done
unset FGB_MODELS_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:circles_move -s 10 -c 1 -o gpurun_out/r2i/move_new_step10 python tools/run_circles.py --steps 12 --iter-mode -1 > gpurun_out/r2i/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:circles_move -s 300 -c 1 -o gpurun_out/r2i/move_new_step300 python tools/run_circles.py --steps 302 --iter-mode -1 > gpurun_out/r2i/ncu2.log 2>&1
ls -la gpurun_out/r2i
