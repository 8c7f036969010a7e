#!/bin/bash
# Build differently-configured copies of libss_b200.so into variants/ (git-ignored, travels with gpurun) so that one GPU
# call can time them side by side:   tools/build_variants.sh && gpurun -- 'tools/gpu_variants.sh variants/*.so'
# Every variant listed here is kept functionally verified by tests/test_simt_emu.py (SIMT emulation) -- time it on a B200
# and run `SS_B200_LIB=variants/<lib> python -m pytest tests -m gpu -q` before making it the default.
# BASE=<git rev> additionally builds that revision's sources as variants/libss_base.so (same-box A/B against history).
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I include"
SRC=steppingstone_b200/csrc
nvcc $FLAGS -I $SRC                    -o variants/libss_default.so $SRC/ss_kernels.cu &
nvcc $FLAGS -I $SRC -DSS_SPARSE_LDL=0  -o variants/libss_dense_ldl.so $SRC/ss_kernels.cu &
nvcc $FLAGS -I $SRC -DSS_FAST_DIV=0    -o variants/libss_ieee_div.so $SRC/ss_kernels.cu &
for extra in "$@"; do   # e.g. tools/build_variants.sh "-DSS_WPB=16"  ->  variants/libss__DSS_WPB_16.so
  nvcc $FLAGS -I $SRC $extra -o "variants/libss_$(echo "$extra" | tr -c 'A-Za-z0-9\n' '_').so" $SRC/ss_kernels.cu &
done
if [ -n "$BASE" ]; then
  tmp=$(mktemp -d); mkdir -p $tmp/csrc $tmp/include
  git show $BASE:$SRC/ss_kernels.cu > $tmp/csrc/ss_kernels.cu; git show $BASE:$SRC/ss_device.cuh > $tmp/csrc/ss_device.cuh
  git show $BASE:include/ss_env.h > $tmp/include/ss_env.h; git show $BASE:include/ss_model.h > $tmp/include/ss_model.h
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I $tmp/include -I $tmp/csrc -o variants/libss_base.so $tmp/csrc/ss_kernels.cu &
fi
wait
ls -la variants/
