#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: the same sources with extra compile-time flags ->
# foundationpose_b200/lib/variants/libfpose_<name>.so (load with FPOSE_LIB_PATH; A/B experiments, traces)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=foundationpose_b200/lib/variants; mkdir -p $out/obj_$name
for f in foundationpose_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
    -c $f -o $out/obj_$name/$(basename ${f%.cu}).o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out/libfpose_$name.so $out/obj_$name/*.o -lcudart -Xlinker -rpath,/usr/local/cuda/lib64
rm -rf $out/obj_$name
ls -la $out/libfpose_$name.so
