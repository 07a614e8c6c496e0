#!/bin/bash
# usage: build_variant.sh name "nvcc flags"  ->  build/<name>/libcpecan_b200.so (kernel-variant experiments; see tools/run_variants.sh)
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
make -s -f $root/cpecan_b200/csrc/Makefile OUT=$root/build/$name EXTRA_NVFLAGS="$*" $root/build/$name/libcpecan_b200.so
