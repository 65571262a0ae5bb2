#!/bin/bash
O=gpurun_out/r2g
mkdir -p $O
: > $O/sizes.txt
for size in 512x256x256 1024x1024x128 1024x1024x512 1024x1024x1024; do
    B200_TRICUBIC_ROWS=7 timeout 200 bash tools/quick.sh tricubic double $size >> $O/sizes.txt 2>> $O/err.txt
done
STEPS=40 B200_TRICUBIC_ROWS=7 timeout 200 bash tools/quick.sh tricubic double 512x256x256 >> $O/sizes.txt 2>> $O/err.txt
STEPS=2 B200_TRICUBIC_ROWS=7 timeout 200 bash tools/quick.sh tricubic double 1024x1024x1024 >> $O/sizes.txt 2>> $O/err.txt
cat $O/sizes.txt
