#!/bin/bash
# Builds oracle/_ref/ref_sim_fgb: the reference library (the objects build_ref.sh produced) with four of its method
# bodies replaced by integration/fgb_reference_shims.cu, i.e. FLAME GPU 2's own CUDASimulation::step() running on the
# kernels of libflamegpu2_b200.so.  The reference's objects are not recompiled: the
# symbols the shim redefines are weakened in copies of the four objects that define them, so the shim's strong
# definitions win at link time.  Test infrastructure (tests/test_dropin_gpu.py compares ref_sim_fgb with ref_sim).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
REF=${FGB_REFERENCE:-/root/reference}
OUT="$HERE/../_ref"
OBJ="$OUT/obj"
DOBJ="$OUT/obj_dropin"
if [ ! -d "$REF/src/flamegpu" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref"; exit 0; fi
if [ ! -f "$OBJ/src_flamegpu_simulation_detail_CUDAScatter.cu.o" ]; then echo "run build_ref.sh first"; exit 1; fi
if [ ! -f "$ROOT/flamegpu2_b200/lib/libflamegpu2_b200.so" ]; then echo "build libflamegpu2_b200.so first"; exit 1; fi
JSON=$(python - <<'PY'
import sysconfig, os
print(os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty"))
PY
)
mkdir -p "$DOBJ"
FLAGS="-x cu -rdc=true --expt-relaxed-constexpr -std=c++20 -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo \
 -I$HERE/shim -I$REF/include -I$JSON -I$ROOT/include -I$ROOT -DFLAMEGPU_SEATBELTS=0 -DFLAMEGPU_TELEMETRY_SUPPRESS_NOTICE \
 -include fstream -include sstream -w"
# object -> pattern of the (mangled) symbols the shim takes over
weaken() {
  src="$OBJ/$1"; dst="$DOBJ/$1"; pat="$2"
  args=""
  for s in $(nm "$src" | awk '$2 == "T" {print $3}' | grep -E "$pat"); do args="$args --weaken-symbol=$s"; done
  if [ -z "$args" ]; then echo "no symbol matches $pat in $1"; exit 1; fi
  objcopy $args "$src" "$dst"
  echo "weakened in $1:$(echo $args | sed 's/--weaken-symbol=/\n   /g')"
}
weaken src_flamegpu_runtime_messaging_MessageSpatial3D.cu.o '^_ZN8flamegpu16MessageSpatial3D16CUDAModelHandler10buildIndexE'
weaken src_flamegpu_runtime_messaging_MessageSpatial2D.cu.o '^_ZN8flamegpu16MessageSpatial2D16CUDAModelHandler10buildIndexE'
weaken src_flamegpu_runtime_messaging_MessageBucket.cu.o '^_ZN8flamegpu13MessageBucket16CUDAModelHandler10buildIndexE'
weaken src_flamegpu_simulation_detail_CUDAScatter.cu.o '^_ZN8flamegpu6detail11CUDAScatter7scatterEj'
nvcc $FLAGS -c "$ROOT/integration/fgb_reference_shims.cu" -o "$DOBJ/fgb_reference_shims.o"
nvcc $FLAGS -c "$HERE/ref_sim.cu" -o "$DOBJ/ref_sim.o"
OBJS=$(ls "$OBJ"/*.o | grep -v -e MessageSpatial3D.cu.o -e MessageSpatial2D.cu.o -e MessageBucket.cu.o -e detail_CUDAScatter.cu.o)
nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true "$DOBJ/ref_sim.o" $OBJS "$DOBJ"/src_*.o "$DOBJ/fgb_reference_shims.o" \
  -o "$OUT/ref_sim_fgb" -lcuda -lnvrtc -L"$ROOT/flamegpu2_b200/lib" -lflamegpu2_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../flamegpu2_b200/lib'
rm -f "$DOBJ/ref_sim.o"
echo "built $OUT/ref_sim_fgb"
