#!/bin/bash
# Builds the UNMODIFIED reference library (FLAME GPU 2, /root/reference) for sm_100a into
# oracle/_ref/ so that the GPU box can run the reference's own CUDASimulation::step()
# beside ours.  Test/bench infrastructure only; nothing here is on the product path.
#
# The reference's CMake build needs network (FetchContent of CCCL, Jitify2, tinyxml2,
# nlohmann_json), so the sources are compiled where they lie with plain nvcc, leaving out
# detail/JitifyCache.cu, io/XML*.cu and the MPI TUs; a ~20 line jitify shim and one stub TU
# close the link.  Outputs: oracle/_ref/libflamegpu_ref.a and
# oracle/_ref/ref_sim (driver in ref_sim.cu).  ~10 min on 8 cores from scratch.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${FGB_REFERENCE:-/root/reference}
OUT="$HERE/../_ref"
OBJ="$OUT/obj"
JSON=$(python - <<'PY'
import sysconfig, os
p = os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty")
print(p)
PY
)
if [ ! -d "$REF/src/flamegpu" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref"; exit 0; fi
if [ ! -f "$JSON/nlohmann/json.hpp" ]; then echo "nlohmann/json.hpp not found under $JSON"; exit 1; fi
mkdir -p "$OBJ"
FLAGS="-x cu -rdc=true --expt-relaxed-constexpr -std=c++20 -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo \
 -I$HERE/shim -I$REF/include -I$JSON -DFLAMEGPU_SEATBELTS=0 -DFLAMEGPU_TELEMETRY_SUPPRESS_NOTICE \
 -include fstream -include sstream -w"
compile_one() {
  f="$1"; o="$OBJ/$(echo "$f" | sed "s#$REF/##; s#/#_#g").o"
  if [ -f "$o" ] && [ "$o" -nt "$f" ]; then return 0; fi
  nvcc $FLAGS -c "$f" -o "$o" > "$o.log" 2>&1 || { echo "FAIL $f"; tail -5 "$o.log"; return 1; }
  echo "OK   $f"
}
export -f compile_one; export OBJ REF FLAGS
find "$REF/src/flamegpu" \( -name '*.cu' -o -name '*.cpp' \) \
  | grep -v -e 'detail/JitifyCache.cu' -e 'io/XMLLogger.cu' -e 'io/XMLStateReader.cu' -e 'io/XMLStateWriter.cu' -e '/MPI' -e '/visualiser/' \
  | sort > "$OUT/tus.txt"
xargs -P ${FGB_JOBS:-8} -I{} bash -c 'compile_one {}' < "$OUT/tus.txt"
nvcc $FLAGS -c "$HERE/link_stubs.cu" -o "$OBJ/link_stubs.o"
rm -f "$OUT/libflamegpu_ref.a"
ar rcs "$OUT/libflamegpu_ref.a" "$OBJ"/*.o
nvcc $FLAGS -c "$HERE/ref_sim.cu" -o "$OUT/ref_sim.o"
nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true "$OUT/ref_sim.o" "$OBJ"/*.o -o "$OUT/ref_sim" -lcuda -lnvrtc
rm -f "$OUT/ref_sim.o"
echo "built $OUT/ref_sim"
