#!/bin/bash
# usage: tools/ncu_one.sh <test> <real> <size> <tag>   -> gpurun_out/prof_<test>_<real>_<tag>.ncu-rep (one launch, full set)
t=$1; r=$2; s=$3; tag=$4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 1 -f \
  -o gpurun_out/prof_${t}_${r}_${tag} python bench.py --test $t --real $r --size $s --steps 2 --warmup 3 --suite none --no-e2e --no-cpu > gpurun_out/ncu_${t}_${r}_${tag}.log 2>&1
