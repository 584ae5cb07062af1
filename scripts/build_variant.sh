#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: compiles csrc/*.cu with extra defines into supernormal_b200/lib/variants/libsnb200_NAME.so
# (experiments: swap it over lib/libsnb200.so on the GPU box, measure, restore)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=supernormal_b200/lib/variants; obj=supernormal_b200/build/variant_$name
mkdir -p $out $obj
for f in supernormal_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $obj/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -o $out/libsnb200_$name.so $obj/*.o -lcudart
echo built $out/libsnb200_$name.so
