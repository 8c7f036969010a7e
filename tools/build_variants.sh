#!/bin/bash
# Build differently-configured copies of libss_b200.so into variants/ (git-ignored, travels with gpurun) so that one GPU
# call can time them side by side:   tools/build_variants.sh && gpurun -- 'tools/gpu_variants.sh variants/*.so'
# Every variant listed here is kept functionally verified by tests/test_simt_emu.py (SIMT emulation) -- time it on a B200
# and run `SS_B200_LIB=variants/<lib> python -m pytest tests -m gpu -q` before making it the default.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I include -I steppingstone_b200/csrc"
nvcc $FLAGS                    -o variants/libss_default.so steppingstone_b200/csrc/ss_kernels.cu
nvcc $FLAGS -DSS_SPARSE_LDL=1  -o variants/libss_sparse_ldl.so steppingstone_b200/csrc/ss_kernels.cu
ls -la variants/
