#!/bin/bash
# tuning builds of the tracking kernel: variants/libcomo_b200_<name>.so (select with COMO_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
python build.py >/dev/null
mkdir -p variants build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
OTHERS=$(ls build/*.o | grep -v "/track.o" | grep -v "track_v")
build() {  # name, flags...
  name=$1; shift
  nvcc $F "$@" -c como_b200/csrc/track.cu -o build/track_v_$name.o
  nvcc -shared -o variants/libcomo_b200_$name.so build/track_v_$name.o $OTHERS -lcudart
  echo "built variants/libcomo_b200_$name.so ($*)"
}
build s2 -DTRK_STAGES1=2 &
build o3 -DTRK_MAX_OCC=3 &
wait
