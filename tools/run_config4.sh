#!/bin/bash
# tools/run_config4.sh [tag] -- run ON AN 8-GPU BOX (gpurun --gpus 8): BASELINE configs[3],
# laplacian + wave13pt on a 2048^3 double grid, niters=100, cut into z-slabs over 2 / 4 / 8 B200s
# (strong scaling of the fixed 2048^3 grid; fused halo push over NVLink).  One JSON line per run
# in gpurun_out/config4_<tag>_<test>_n<N>.json.
tag=${1:-r1}
mkdir -p gpurun_out
port=29700
for n in 8 4 2; do
  for t in laplacian wave13pt; do
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --test $t --real double --size 2048x2048x$((2048/n)) --niters 100 --steps 2 --warmup 3 \
      --no-e2e --no-cpu --suite none > gpurun_out/config4_${tag}_${t}_n${n}.json 2> gpurun_out/config4_${tag}_${t}_n${n}.err
    tail -c 600 gpurun_out/config4_${tag}_${t}_n${n}.json | cut -c1-400
  done
done
