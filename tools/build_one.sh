#!/bin/bash
# rebuild one system unit (default CARTPOLE) and relink the library in place: fast iteration on engine.cuh
sys=${1:-CARTPOLE}; cls=${2:-SysCartpole}; shift; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fopenmp -DMYR_SYS_CLASS=$cls "$@" -c myriad_b200/csrc/sys_unit.cu -o build/sys_$sys.o 2>&1 | grep -i "error\|warning.*spill" -A3
nvcc -shared -o myriad_b200/libmyriad_b200.so build/sys_*.o build/node_*.o build/api.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -lgomp
cuobjdump --dump-resource-usage build/sys_$sys.o 2>/dev/null | grep -A1 "Function" | paste - - | grep "ipm_kernelINS_9Trap\|ipm_kernelINS_14Herm" | sed 's/_ZN3myr//' | cut -c1-150
