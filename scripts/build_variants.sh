#!/bin/bash
# Build experiment variants of the library: scripts/variants/lib_<name>.so, selected at run time with DQ_LIB_PATH.
# usage: scripts/build_variants.sh name1="-DFLAG=1 -DOTHER" name2="..."
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/variants/obj
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  objs=""
  for f in admm_fwd admm_fwd_tpp qp_bwd qcqp_bwd boxqp_bwd large_n api; do
    o=scripts/variants/obj/${name}_$f.o
    nvcc $flags -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c diffqcqp_b200/csrc/$f.cu -o $o &
    objs="$objs $o"
  done
  wait
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scripts/variants/lib_$name.so $objs
  echo built scripts/variants/lib_$name.so
done
