#!/bin/bash
# Builds an A/B variant of the library with extra -D flags:  tools/build_variant.sh <name> [-DFLAG ...]  -> build/variants/<name>/libsdm_b200.so
# ALL=1 recompiles every source with the flags (default: only the two tcgen05 kernels).
# Select it at run time with SDM_B200_LIB=build/variants/<name>/libsdm_b200.so
set -e
name=$1; shift
out=build/variants/$name
mkdir -p $out
for f in simple-diffusion-model_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$ALL" == "1" ] || [ "$b" == "gemm_tn" ] || [ "$b" == "igemm_nt" ] || [ ! -f build/$b.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Isimple-diffusion-model_b200/csrc -Iinclude -Xcompiler -fPIC "$@" -c $f -o $out/$b.o &
  else
    cp build/$b.o $out/$b.o
  fi
done
wait
nvcc -shared -o $out/libsdm_b200.so $out/*.o -lcudart
echo built $out/libsdm_b200.so
