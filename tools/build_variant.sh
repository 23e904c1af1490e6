#!/bin/bash
# tools/build_variant.sh <tag> <extra nvcc flags...>: CARTPOLE unit rebuilt with extra flags, linked with the other objects into build/lib_<tag>.so
tag=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DMYR_SYS_CLASS=SysCartpole "$@" -c myriad_b200/csrc/sys_unit.cu -o build/var_${tag}_CARTPOLE.o 2>&1 | grep -i "error" -A3; test -f build/var_${tag}_CARTPOLE.o || exit 1
objs=$(ls build/sys_*.o build/node_*.o build/api.o | grep -v "sys_CARTPOLE.o")
nvcc -shared -o build/lib_${tag}.so build/var_${tag}_CARTPOLE.o $objs -gencode arch=compute_100a,code=sm_100a
cuobjdump --dump-resource-usage build/var_${tag}_CARTPOLE.o 2>/dev/null | grep -A1 "Function" | paste - - | grep "ipm_kernelINS_9Trap" | sed 's/_ZN3myr//' | cut -c1-150
