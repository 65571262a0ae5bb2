#!/bin/bash
# tools/r2f_run.sh -- ncu --set full of the tricubic double kernel, form $FORM, at 512x256x256 and 1024x1024x128
O=gpurun_out/r2f
mkdir -p $O
for size in 512x256x256 1024x1024x128; do
B200_TRICUBIC_ROWS=${FORM:-7} timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 1 -f \
  -o $O/prof_tricubic_double_f${FORM:-7}_$size python bench.py --test tricubic --real double --size $size --steps 2 --warmup 3 \
  --suite none --no-e2e --no-cpu > $O/ncu_$size.log 2>&1
echo "ncu $size rc=$?"
done
ls -la $O
