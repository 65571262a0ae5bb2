#!/bin/bash
# tools/r1v_run.sh -- run ON THE GPU BOX (gpurun), 1 GPU: final state of round 1 -- smoke(), one ncu --set full capture of the
# default tricubic double kernel, then the whole GPU test suite.
O=gpurun_out/r1v
mkdir -p $O
T0=$(date +%s)
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 120 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 1 -f \
  -o $O/prof_tricubic_double_default python bench.py --test tricubic --real double --size 512x256x256 --steps 2 --warmup 3 \
  --suite none --no-e2e --no-cpu > $O/ncu.log 2>&1
echo "ncu rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout ${PYTEST_LIMIT:-230} python -m pytest tests -q -m gpu -x --durations=6 > $O/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -12 $O/pytest_gpu.log
