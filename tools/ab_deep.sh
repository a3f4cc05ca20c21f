#!/bin/bash
# Same-box A/B of compile-time variants of the cluster-split kernel: builds bc_mpc_b200/_ab/lib_<name>.so for every
# "name:flags" argument from the objects of the regular build (only the tm_deep*.cu translation units - all of them - are recompiled).
#   tools/ab_deep.sh base: fp32x:-DBCMPC_DEEP_FP32_EXCHANGE st7:-DBCMPC_DEEP_STAGES=7
# then, on the GPU box:  for l in bc_mpc_b200/_ab/lib_*.so; do BCMPC_LIB=$l python tools/deep_bench.py; done
set -e
cd "$(dirname "$0")/../bc_mpc_b200"
mkdir -p _ab
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"; [ "$flags" = "$spec" ] && flags=""
  dir=_ab/obj_$name; mkdir -p $dir
  for f in $(ls csrc/tm_deep*.cu | xargs -n1 basename | sed 's/\.cu$//'); do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC $flags -c -o $dir/$f.o csrc/$f.cu &
  done
  wait
  objs=$(ls _obj/*.o | grep -v "/tm_deep")
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -o _ab/lib_$name.so $objs $dir/*.o
  echo built _ab/lib_$name.so "($flags)"
done
