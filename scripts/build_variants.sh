#!/bin/bash
# dev tool: build alternative libvisma_b200 variants for scripts/ab_pass.py
#   scripts/build_variants.sh name1 "DEFS1" name2 "DEFS2" ...  -> build/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  make -s -C visma_b200/csrc -j8 OUT=$PWD/build/variants/lib_$name.so OBJDIR=$PWD/build/variants/obj_$name DEFS="$defs" > /dev/null
  grep -A2 "k_passILi1" build/variants/obj_$name/icp.ptxas.log | grep -E "registers|spill" | sed "s/^/$name: /"
done
