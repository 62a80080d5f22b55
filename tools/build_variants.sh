#!/bin/bash
# build tuning variants of the native libraries into lambrex_b200/_lib_<name>/ (select with LBX_LIB_DIR=_lib_<name>)
set -e
cd "$(dirname "$0")/../lambrex_b200"
build() {   # name, RO flags
  make -s -C csrc OUT=../_lib_$1 RO="$2" 2>&1 | grep -i "error" || true
  make -s -C host OUT=../_lib_$1 ../_lib_$1/liblambrex.so 2>&1 | grep -i "error" || true
  grep -A3 "k_mf_cs_rows" _lib_$1/kernels_fast.ptxas.txt | grep -i "registers\|spill" | head -4
}
build t128c6 "-DLBX_RO_THREADS=128 -DLBX_RO_MIN_CTAS=6"
build t256c4 "-DLBX_RO_THREADS=256 -DLBX_RO_MIN_CTAS=4"
build t128c8 "-DLBX_RO_THREADS=128 -DLBX_RO_MIN_CTAS=8"
build t256c2 "-DLBX_RO_THREADS=256 -DLBX_RO_MIN_CTAS=2"
