#!/bin/bash
# tools/r1t_run.sh -- run ON THE GPU BOX (gpurun), 1 GPU: tricubic with the transient ring (B200_TRICUBIC_ROWS=4) against
# the default (2), parity of every form, the bench line + launch list + DRAM traffic of the headline, then the GPU parity tests.
O=gpurun_out/r1t
mkdir -p $O
T0=$(date +%s)
left() { echo $(( ${DEADLINE:-420} - ($(date +%s) - T0) )); }
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k tricubic_row_variants > $O/variants.log 2>&1
echo "variants rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
: > $O/ab.txt
for cfg in "double 512x256x256" "double 1024x1024x512" "double 1024x1024x1024" "float 512x256x256" "float 1024x1024x512"; do
  for rows in 2 4; do
    echo -n "rows=$rows " >> $O/ab.txt
    B200_TRICUBIC_ROWS=$rows timeout 150 bash tools/quick.sh tricubic $cfg >> $O/ab.txt 2>> $O/ab.err || echo "FAILED $cfg rows=$rows" >> $O/ab.txt
  done
done
cat $O/ab.txt
echo "ab t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
eval $(python - <<'PY'
import re
best = {}
for line in open("gpurun_out/r1t/ab.txt"):
    m = re.match(r"rows=(\d) tricubic \S+ (double|float) niters=\d+ ([\d.]+) frac", line)
    if m:
        rows, real, v = m.groups()
        best.setdefault(real, {}).setdefault(rows, []).append(float(v))
for real, env in (("double", "F64"), ("float", "F32")):
    d = best.get(real, {})
    ok = "4" in d and "2" in d and len(d["4"]) == len(d["2"]) and all(b > 1.015 * a for a, b in zip(d["2"], d["4"]))
    print(f"export B200_TRICUBIC_ROWS_{env}={'4' if ok else '2'};")
PY
)
echo "winners: F64=$B200_TRICUBIC_ROWS_F64 F32=$B200_TRICUBIC_ROWS_F32" | tee $O/winners.txt
timeout 240 python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 400 --csv \
    --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --suite none --no-cpu > /dev/null 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:stream_kernel -s 30 -c 20 --csv --log-file $O/traffic.csv \
    python bench.py --steps 5 --warmup 3 --suite none --no-e2e --no-cpu > /dev/null 2>&1
echo "ncu lists t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
if [ $(left) -gt 60 ]; then
  timeout $(( $(left) - 10 )) python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x --durations=8 > $O/pytest_parity.log 2>&1
  echo "pytest parity+fullsize rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
  tail -14 $O/pytest_parity.log
fi
