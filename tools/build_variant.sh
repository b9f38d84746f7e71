#!/bin/bash
# Build a named variant of libeoe_b200.so from a copy of csrc/:
#   tools/build_variant.sh <name> [<replacement gemm header> | ""] [extra nvcc flags, e.g. -DEOE_F16_GELU_EXACT=1]
# The variant is selected at run time with EOE_B200_LIB=tools/_variants/libeoe_b200_<name>.so (A/B timing, eoe_b200/_lib.py).
set -e
name=$1; hdr=$2; shift; shift || true
d=/tmp/eoe_variant_$name; rm -rf $d; mkdir -p $d/eoe_b200/csrc $d/include tools/_variants
cp eoe_b200/csrc/* $d/eoe_b200/csrc/; cp include/eoe_b200.h $d/include/
[ -n "$hdr" ] && cp $hdr $d/eoe_b200/csrc/gemm_sm100.cuh
objs=""
for f in $d/eoe_b200/csrc/*.cu; do
  o=${f%.cu}.o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "-DEOE_BUILD_ID=\"variant-$name\"" "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -o tools/_variants/libeoe_b200_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcuda
ls -la tools/_variants/libeoe_b200_$name.so
