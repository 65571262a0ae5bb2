#!/bin/bash
# tools/r1s_run.sh -- run ON THE GPU BOX (gpurun), 1 GPU: tricubic with 8 vs 10 consumer warps (B200_TRICUBIC_ROWS=2|3),
# parity of every form, the driver tests (incl. the parallel rand() fill), one ncu capture of the 10-warp form.
O=gpurun_out/r1s
mkdir -p $O
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k tricubic_row_variants > $O/variants.log 2>&1
echo "variants rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
: > $O/ab.txt
for real in double float; do
  for size in 512x256x256 1024x1024x512; do
    for rows in 2 3; do
      echo -n "rows=$rows " >> $O/ab.txt
      B200_TRICUBIC_ROWS=$rows timeout 120 bash tools/quick.sh tricubic $real $size >> $O/ab.txt 2>> $O/ab.err || echo "FAILED $real $size rows=$rows" >> $O/ab.txt
    done
  done
done
for rows in 2 3; do
  echo -n "rows=$rows " >> $O/ab.txt
  B200_TRICUBIC_ROWS=$rows timeout 150 bash tools/quick.sh tricubic double 1024x1024x1024 >> $O/ab.txt 2>> $O/ab.err || echo "FAILED double 1024^3 rows=$rows" >> $O/ab.txt
done
cat $O/ab.txt
echo "ab t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
B200_TRICUBIC_ROWS=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 1 -f \
  -o $O/prof_tricubic_double_rows3 python bench.py --test tricubic --real double --size 512x256x256 --steps 2 --warmup 3 \
  --suite none --no-e2e --no-cpu > $O/ncu_rows3.log 2>&1
echo "ncu t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 400 python -m pytest tests/test_gpu_drivers.py -q -m gpu -x > $O/pytest_drivers.log 2>&1
echo "pytest drivers rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -3 $O/pytest_drivers.log
