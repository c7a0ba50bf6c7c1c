#!/bin/bash
# Build an instrumented variant of the library next to the product one:
#   tools/build_variant.sh trace -DKPD_WS_TRACE        -> keypoint_diffusion_b200/libkpdiff_b200_trace.so
# Use it with KPD_LIB=keypoint_diffusion_b200/libkpdiff_b200_<name>.so python tools/...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/keypoint_diffusion_b200/csrc
out=$root/build/$name
mkdir -p "$out"
for f in row_ops graph_build egnn gvp ddpm_step sampler tc_gemm; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c "$src/$f.cu" -o "$out/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/keypoint_diffusion_b200/libkpdiff_b200_$name.so" "$out"/*.o
echo built "keypoint_diffusion_b200/libkpdiff_b200_$name.so"
