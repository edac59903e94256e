#!/bin/bash
# development aid: builds liboak_<name>.so with extra -D flags into scripts/ubench/
# usage: scripts/build_variants.sh idx1:-DOAK_IDX_MODE=1 idx2:-DOAK_IDX_MODE=2 ...
set -e
cd "$(dirname "$0")/../orthogonal-additive-gaussian-processes_b200/csrc"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags \
      -I../../include -I. -shared -o ../../scripts/ubench/liboak_$name.so oak_*.cu \
      -L/usr/local/cuda/lib64 -lcublas -lcusolver -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 ) &
done
wait
