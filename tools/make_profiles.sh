#!/bin/bash
# tools/make_profiles.sh <tag>  -- run ON THE GPU BOX (gpurun).  Produces under gpurun_out/:
#   launches_<tag>.csv        ncu launch list of the default bench command (steady state: --cache-control none)
#   prof_wave13pt_<tag>.ncu-rep  one `--set full` capture of the headline kernel (ncu's default cache flush)
#   traffic_<tag>.csv         dram bytes per launch of the headline kernel, steady state
#   bench_<tag>.json          the bench line itself (NOT under a profiler)
tag=${1:-r1}
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 400 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 20 --warmup 3 --suite none --no-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:stream_kernel -s 30 -c 20 --csv --log-file gpurun_out/traffic_${tag}.csv \
    python bench.py --steps 5 --warmup 3 --suite none --no-e2e --no-cpu > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 30 -c 1 -f \
    -o gpurun_out/prof_wave13pt_${tag} python bench.py --steps 3 --warmup 3 --suite none --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out | grep ${tag}
# one `--set full` capture per stencil kernel (double, README size) -> tools/collect_profiles.py writes <tag>_ncu_all_kernels.txt
if [ "${ALL_KERNELS:-1}" = "1" ]; then
  for t in laplacian divergence gradient uxx1 lapgsrb tricubic; do bash tools/ncu_one.sh $t double 512x256x256 $tag; done
  for t in jacobi gaussblur gameoflife; do bash tools/ncu_one.sh $t double 512x65536x1 $tag; done
  bash tools/ncu_one.sh gameoflife float 512x65536x1 $tag
  bash tools/ncu_one.sh lapgsrb float 512x256x256 $tag
  bash tools/ncu_one.sh tricubic float 512x256x256 $tag
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:matmul -c 4 -f -o gpurun_out/prof_matmul_$tag python tools/matmul_bench.py 4096 1 > gpurun_out/ncu_matmul_$tag.log 2>&1
fi
timeout 600 python tools/ref_cuda_table.py gpurun_out/ref_cuda_$tag.json > gpurun_out/ref_cuda_$tag.log 2>&1
ls -la gpurun_out | grep ${tag}
