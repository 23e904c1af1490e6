#!/bin/bash
# tools/build_variant.sh <tag> <extra nvcc flags...>: single-system library build/lib_<tag>.so with extra flags (A/B runs via
# MYR_LIB).  The system is CARTPOLE unless VAR_SYS names another generated struct (e.g. VAR_SYS=SysSeir).
tag=$1; shift
SYS=${VAR_SYS:-SysCartpole}
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fopenmp"
nvcc $F -DMYR_SYS_CLASS=$SYS "$@" -c myriad_b200/csrc/sys_unit.cu -o build/var_${tag}_$SYS.o 2>&1 | grep -i "error" -A3; test -f build/var_${tag}_$SYS.o || exit 1
api=build/api_only_$SYS.o
if [ ! -f $api ] || [ myriad_b200/csrc/api.cu -nt $api ] || [ include/myriad_b200.h -nt $api ]; then
  nvcc $F "-DMYR_BUILD_SYSTEMS(X)=X($SYS)" "-DMYR_BUILD_NODE_SYSTEMS(X)=" -c myriad_b200/csrc/api.cu -o $api
fi
nvcc -shared -o build/lib_${tag}.so build/var_${tag}_$SYS.o $api -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -lgomp
cuobjdump --dump-resource-usage build/var_${tag}_$SYS.o 2>/dev/null | grep -A1 "Function" | paste - - | grep "ipm_kernelINS_9Trap\|ipm_kernelINS_14Herm" | sed 's/_ZN3myr//' | cut -c1-150
