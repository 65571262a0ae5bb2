#!/bin/bash
# tools/r2_first_call.sh -- run ON THE GPU BOX (gpurun), 1 GPU, ~4 minutes: the first call of the next round.
#   1. parity run of the opt-in kernel forms (tests/test_gpu_experimental.py),
#   2. A/B of the float tile policy (B200_TILE_POLICY=0 default form, 1 model's choice) at C1 and C2,
#   3. driver t_load / t_save with B200_PINNED_HOST.
# Everything lands in gpurun_out/r2a/.  If 1 passes and 2 shows the C1 gain without a C2 loss, make policy 1 the default
# (b200_launch.cuh: tile_policy) and move the test into tests/test_gpu_parity.py.
O=gpurun_out/r2a
mkdir -p $O
T0=$(date +%s)
B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu -s > $O/experimental.log 2>&1
echo "experimental rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
: > $O/tile_policy.txt
for t in laplacian wave13pt divergence gradient lapgsrb; do
  for size in 512x256x256 1024x1024x512; do
    for pol in 0 1; do
      echo -n "policy=$pol " >> $O/tile_policy.txt
      B200_TILE_POLICY=$pol timeout 120 bash tools/quick.sh $t float $size >> $O/tile_policy.txt 2>> $O/tile_policy.err || echo "FAILED $t $size policy=$pol" >> $O/tile_policy.txt
    done
  done
done
cat $O/tile_policy.txt
echo "tile policy t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
