#!/bin/bash
# tools/r2b_run.sh -- on the GPU box: full gpu test suite + the default bench (N=1) + reference arm
O=gpurun_out/r2b
mkdir -p $O
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -5 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
echo "ref rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
python tools/suite_table.py $O/bench.json > $O/suite_table.md 2>/dev/null
head -c 1500 $O/bench.json; echo; cat $O/bench_ref.json | head -c 600
