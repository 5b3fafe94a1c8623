#!/bin/bash
# Kernel-tuning build: recompile ONE source of the library with extra -D flags and link it with the stock objects.
#   tools/build_variant.sh <name> <source stem> <nvcc -D flags...>   ->  variants/libaptp_<name>.so  (use with APTP_LIB=...)
set -e
name=$1; stem=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p variants
python -m diffusion_pruning_b200.build > /dev/null
objs=""
for o in diffusion_pruning_b200/build/*.o; do
  [ "$(basename $o .o)" == "$stem" ] || objs="$objs $o"
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  "$@" -c diffusion_pruning_b200/csrc/$stem.cu -o variants/${stem}_$name.o
/usr/local/cuda/bin/nvcc -shared -o variants/libaptp_$name.so $objs variants/${stem}_$name.o -lcudart 2>/dev/null
echo variants/libaptp_$name.so
