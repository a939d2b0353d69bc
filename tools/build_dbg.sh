#!/bin/sh
# profiling build of the library with the cycle counters compiled in (never the product build)
cd "$(dirname "$0")/.." && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -DSTDADK_PF_DEBUG -o st_dadk_b200/libstdadk_dbg.so st_dadk_b200/csrc/api.cu
