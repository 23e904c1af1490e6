#!/bin/bash
timeout 100 python tools/k1_bench.py 2>&1 | grep trap
for v in "$@"; do echo "variant $v"; MYR_LIB=$PWD/build/lib_$v.so timeout 100 python tools/k1_bench.py 2>&1 | grep trap; done
