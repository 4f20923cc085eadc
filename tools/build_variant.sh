#!/bin/bash
# Builds an experimental fp16 variant of the library + its standalone GEMM check with extra -D flags:
#   tools/build_variant.sh rs4 -DSPG_RES_SLOTS=4   ->  spegnet_b200/csrc/build/variants/rs4/{libspegnet_b200_fp16.so,test_gemm}
set -e
name=$1; shift
cd "$(dirname "$0")/../spegnet_b200/csrc"
out=build/variants/$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in *.cu; do
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DSPG_FP16 "$@" -c $f -o $out/${f%.cu}.o &
done
wait
nvcc $ARCH -shared -o $out/libspegnet_b200_fp16.so $out/*.o -cudart shared
nvcc $ARCH -O2 -std=c++17 -DSPG_FP16 "$@" -o $out/test_gemm ../../tests/cuda/test_gemm.cu -L$out -lspegnet_b200_fp16 -Xlinker -rpath -Xlinker '$ORIGIN' -cudart shared
echo built $out
