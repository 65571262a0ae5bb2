#!/bin/bash
# tools/r2c_run.sh -- tricubic forms A/B: parity of forms 5, 6 and timings of forms 4, 5, 6 (double) / 2, 5, 6 (float)
O=gpurun_out/r2c
mkdir -p $O
: > $O/tricubic_ab.txt
for r in 5 6; do
  B200_TRICUBIC_ROWS=$r timeout 300 python tests/tricubic_variant_check.py > $O/parity_$r.json 2> $O/parity_$r.err
  echo "form $r parity rc=$? $(cat $O/parity_$r.json | cut -c1-200)" | tee -a $O/tricubic_ab.txt
done
for size in 512x256x256 1024x1024x1024; do
  for r in 4 5 6; do
    echo -n "rows=$r " >> $O/tricubic_ab.txt
    B200_TRICUBIC_ROWS=$r timeout 200 bash tools/quick.sh tricubic double $size >> $O/tricubic_ab.txt 2>> $O/err.txt
  done
done
for size in 512x256x256 1024x1024x512; do
  for r in 2 5 6; do
    echo -n "rows=$r " >> $O/tricubic_ab.txt
    B200_TRICUBIC_ROWS=$r timeout 200 bash tools/quick.sh tricubic float $size >> $O/tricubic_ab.txt 2>> $O/err.txt
  done
done
cat $O/tricubic_ab.txt
