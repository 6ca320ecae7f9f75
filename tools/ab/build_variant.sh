#!/bin/bash
# Builds an A/B variant of libraiko_kzg.so: tools/ab/build_variant.sh <name> [-DMACRO=VALUE ...]
# Only the units that see the macros are recompiled (AB_TUS, default tu_msm.cu and kzg_ctx.cu); the
# other objects come from the regular build (run `make -C raiko_b200/csrc` first).  Output:
# raiko_b200/ab/libraiko_kzg_<name>.so (git-ignored, travels with gpurun); select it at run time with
# RAIKO_KZG_LIB_OVERRIDE.  ptxas statistics of k_msm_affine go to raiko_b200/ab/<name>.ptxas.txt.
set -eu
name=$1; shift
root=$(cd "$(dirname "$0")/../.." && pwd)
src=$root/raiko_b200/csrc; out=$root/raiko_b200/ab; obj=$out/_obj_$name
mkdir -p "$obj"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
# AB_TUS: the translation units that see the macros (default: the MSM kernels and the host file)
tus=${AB_TUS:-"tu_msm kzg_ctx"}
for tu in $tus; do ( cd "$src" && nvcc $FLAGS "$@" -c -o "$obj/$tu.o" $tu.cu 2> "$obj/$tu.log" ) & done
wait
objs=""
for tu in kzg_ctx tu_msm tu_path tu_table tu_verify tu_pairing; do
  if [ -f "$obj/$tu.o" ]; then objs="$objs $obj/$tu.o"; else objs="$objs $src/_obj/$tu.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libraiko_kzg_$name.so" $objs
[ -f "$obj/tu_msm.log" ] || touch "$obj/tu_msm.log"
grep -A3 "k_msm_affine" "$obj/tu_msm.log" | grep -E "spill|registers" | sed 's/ptxas info    : //' | tr '\n' ' ' > "$out/$name.ptxas.txt"
echo "$name: $* :: $(cat "$out/$name.ptxas.txt")"
