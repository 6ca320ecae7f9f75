#!/bin/bash
# Builds an A/B variant of libraiko_kzg.so: tools/ab/build_variant.sh <name> [-DMACRO=VALUE ...]
# Only the units that see the affine-kernel / multiply macros are recompiled (tu_msm.cu, kzg_ctx.cu); the
# other objects come from the regular build (run `make -C raiko_b200/csrc` first).  Output:
# raiko_b200/ab/libraiko_kzg_<name>.so (git-ignored, travels with gpurun); select it at run time with
# RAIKO_KZG_LIB_OVERRIDE.  ptxas statistics of k_msm_affine go to raiko_b200/ab/<name>.ptxas.txt.
set -eu
name=$1; shift
root=$(cd "$(dirname "$0")/../.." && pwd)
src=$root/raiko_b200/csrc; out=$root/raiko_b200/ab; obj=$out/_obj_$name
mkdir -p "$obj"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
( cd "$src" && nvcc $FLAGS "$@" -c -o "$obj/tu_msm.o" tu_msm.cu 2> "$obj/tu_msm.log" ) &
( cd "$src" && nvcc $FLAGS "$@" -c -o "$obj/kzg_ctx.o" kzg_ctx.cu 2> "$obj/kzg_ctx.log" ) &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libraiko_kzg_$name.so" "$obj/kzg_ctx.o" "$obj/tu_msm.o" \
    "$src/_obj/tu_path.o" "$src/_obj/tu_table.o" "$src/_obj/tu_verify.o" "$src/_obj/tu_pairing.o"
grep -A3 "k_msm_affine" "$obj/tu_msm.log" | grep -E "spill|registers" | sed 's/ptxas info    : //' | tr '\n' ' ' > "$out/$name.ptxas.txt"
echo "$name: $* :: $(cat "$out/$name.ptxas.txt")"
