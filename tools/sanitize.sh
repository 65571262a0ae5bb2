#!/bin/bash
# tools/sanitize.sh [ngpus] -- compute-sanitizer over the b200 drivers (what `make check / racecheck / synccheck` of a
# <test>/b200/ directory runs, suite_overlay/install_overlay.sh), every test and both precisions under memcheck, the
# mbarrier / TMA / named-barrier protocols of the streaming engine under racecheck and synccheck; with ngpus = 2 also the
# halo-pushing variants (single process, B200_NGPUS=2).  Logs -> gpurun_out/sanitizer/ (summaries copied to profiles/).
N=${1:-1}
O=gpurun_out/sanitizer
mkdir -p $O
B=kernelgen-perf-tests_b200/drivers/bin
TESTS3D="laplacian wave13pt divergence gradient uxx1 lapgsrb tricubic tricubic2 vecadd sincos"
TESTS2D="jacobi gaussblur gameoflife matvec"
run() {  # tool test real args... ; env NG
  tool=$1; t=$2; r=$3; shift 3
  log=$O/${tool}_${t}_${r}_n${NG:-1}.log
  B200_NGPUS=${NG:-1} timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 5 $B/${t}_$r "$@" > $log 2>&1
  rc=$?
  echo "$tool $t $r ngpus=${NG:-1} args=$* rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $O/summary.txt
}
: > $O/summary.txt
for r in double float; do
  for t in $TESTS3D; do run memcheck $t $r 132 37 29 3; done
  for t in $TESTS2D; do run memcheck $t $r 132 301 3; done
  run memcheck matmul $r 130 67 140 2
done
for t in laplacian wave13pt lapgsrb uxx1 tricubic divergence; do run racecheck $t double 132 37 29 2; run synccheck $t double 132 37 29 2; done
for t in jacobi gameoflife; do run racecheck $t double 132 301 2; run synccheck $t double 132 301 2; done
run racecheck wave13pt float 132 37 29 2
run racecheck matmul double 130 67 140 1
# the tcgen05 float GEMM (TMA bulk copies, mbarriers, TMEM): two tiles in n, a ragged K chunk, 2 sweeps
run memcheck matmul float 132 300 260 2
run racecheck matmul float 132 300 260 1
run synccheck matmul float 132 300 260 1
# scalar-loader path (odd pitch: no TMA)
run memcheck laplacian double 63 31 29 2
run racecheck laplacian double 63 31 29 2
# fused two-sweep kernel
B200_FUSE=1 run memcheck jacobi double 260 301 6
B200_FUSE=1 run racecheck jacobi double 260 301 6
if [ "$N" -ge 2 ]; then
  for t in laplacian wave13pt lapgsrb tricubic; do NG=2 run memcheck $t double 132 37 64 3; done
  for t in jacobi gaussblur gameoflife; do NG=2 run memcheck $t double 132 600 3; done
  NG=2 run racecheck laplacian double 132 37 64 3
  NG=2 run racecheck jacobi double 132 600 3
  NG=2 run synccheck wave13pt double 132 37 64 3
fi
echo "---"; grep -c "rc=0" $O/summary.txt; grep -v "rc=0" $O/summary.txt
