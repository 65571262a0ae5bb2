#!/bin/bash
# tools/r2l_run.sh -- 2 GPUs, single process (B200_NGPUS): is the halo-pushing kernel variant itself slower than the plain one?
O=gpurun_out/${OUT:-r2l}
mkdir -p $O
B=kernelgen-perf-tests_b200/drivers/bin
for t in lapgsrb laplacian wave13pt tricubic gameoflife gaussblur jacobi; do
  for real in double float; do
    if [ $t = gameoflife -o $t = gaussblur -o $t = jacobi ]; then a1="512 65536 10"; a2="512 131072 10"; else a1="512 256 256 10"; a2="512 256 512 10"; fi
    echo -n "$t $real N=1: " ; PROFILING_FNAME=$t B200_INIT_THREADS=16 $B/${t}_$real $a1 | grep "kernel time"
    echo -n "$t $real N=2: " ; PROFILING_FNAME=$t B200_INIT_THREADS=16 B200_NGPUS=2 $B/${t}_$real $a2 | grep "kernel time"
  done
done 2>&1 | tee $O/driver_n1_n2.txt
