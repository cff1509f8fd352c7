#!/bin/bash
# Build libltt_b200.so for sm_100a (in-tree; nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
# A/B builds: LTT_OUT=../libltt_b.so LTT_BUILD_DIR=build_b LTT_EXTRA_FLAGS="-DLTT_EPI_WGS=3" ./build.sh
OUT=${LTT_OUT:-../libltt_b200.so}
BD=${LTT_BUILD_DIR:-build}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${LTT_EXTRA_FLAGS}"
mkdir -p $BD
pids=()
for f in *.cu; do
  o=$BD/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o")" ] || [ ../../include/ltt_b200.h -nt "$o" ]; then
    nvcc $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT $BD/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
