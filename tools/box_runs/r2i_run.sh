#!/bin/bash
# tools/r2i_run.sh -- full gpu test suite with durations (incl. the real benchmark on the GPU)
O=gpurun_out/r2i
mkdir -p $O
T0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 ${PYTEST_ARGS} > $O/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -30 $O/pytest_gpu.log
