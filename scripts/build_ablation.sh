#!/bin/bash
# development aid: builds liboak_abl<N>.so with -DOAK_ABLATE=N into scripts/ubench/
set -e
cd "$(dirname "$0")/../orthogonal-additive-gaussian-processes_b200/csrc"
for N in "$@"; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DOAK_ABLATE=$N \
      -I../../include -I. -shared -o ../../scripts/ubench/liboak_abl$N.so oak_*.cu \
      -L/usr/local/cuda/lib64 -lcublas -lcusolver -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 ) &
done
wait
