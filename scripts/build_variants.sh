#!/bin/bash
# Tuning variants of libb200sv.so (compile-time switches of the sweep kernel) for A/B runs in ONE GPU call:
#   B200SV_LIB=qrack_b200/variants/libb200sv_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
mkdir -p qrack_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static"
build() { # name, defines...
  local name=$1; shift
  /usr/local/cuda/bin/nvcc $FLAGS "$@" qrack_b200/csrc/b200sv.cu qrack_b200/csrc/fused.cu -o qrack_b200/variants/libb200sv_$name.so &
}
build noneg -DSV_NO_NEG
build noneg_nopf -DSV_NO_NEG -DSV_NO_PREFETCH
build noneg_nofast -DSV_NO_NEG -DSV_NO_FASTPATH
build nopf -DSV_NO_PREFETCH
wait
ls -la qrack_b200/variants/
