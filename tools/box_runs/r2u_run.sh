#!/bin/bash
# gradient forms A/B in the diagnostics library (B200_GRAD_FORM)
O=gpurun_out/r2u
mkdir -p $O
export B200_LIB=$PWD/kernelgen-perf-tests_b200/libb200stencil_diag.so
: > $O/gradient_forms.txt
for f in 0 1 2 3 4 5 6; do
  for real in double float; do
    for size in 512x256x256 1024x1024x512; do
      echo -n "form=$f " >> $O/gradient_forms.txt
      B200_GRAD_FORM=$f timeout 120 bash tools/quick.sh gradient $real $size >> $O/gradient_forms.txt 2>> $O/err.txt
    done
  done
done
cat $O/gradient_forms.txt
