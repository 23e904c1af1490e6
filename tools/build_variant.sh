#!/bin/bash
# tools/build_variant.sh <tag> <extra nvcc flags...>: CARTPOLE-only library build/lib_<tag>.so with extra flags (A/B runs via MYR_LIB)
tag=$1; shift
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fopenmp"
nvcc $F -DMYR_SYS_CLASS=SysCartpole "$@" -c myriad_b200/csrc/sys_unit.cu -o build/var_${tag}_CARTPOLE.o 2>&1 | grep -i "error" -A3; test -f build/var_${tag}_CARTPOLE.o || exit 1
if [ ! -f build/api_cartonly.o ] || [ myriad_b200/csrc/api.cu -nt build/api_cartonly.o ] || [ include/myriad_b200.h -nt build/api_cartonly.o ]; then
  nvcc $F "-DMYR_BUILD_SYSTEMS(X)=X(SysCartpole)" "-DMYR_BUILD_NODE_SYSTEMS(X)=" -c myriad_b200/csrc/api.cu -o build/api_cartonly.o
fi
nvcc -shared -o build/lib_${tag}.so build/var_${tag}_CARTPOLE.o build/api_cartonly.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -lgomp
cuobjdump --dump-resource-usage build/var_${tag}_CARTPOLE.o 2>/dev/null | grep -A1 "Function" | paste - - | grep "ipm_kernelINS_9Trap\|ipm_kernelINS_14Herm" | sed 's/_ZN3myr//' | cut -c1-150
