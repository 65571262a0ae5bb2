#!/bin/bash
# tools/r1r_run.sh -- run ON THE GPU BOX (gpurun), 1 GPU, ~10 minutes: A/B of the two tricubic kernel forms
# (B200_TRICUBIC_ROWS=1|2), parity of both, one ncu capture of the winner, tricubic GPU tests and the bench line
# under the winning setting.  Everything lands in gpurun_out/r1r/.
O=gpurun_out/r1r
mkdir -p $O
T0=$(date +%s)
left() { echo $(( ${DEADLINE:-660} - ($(date +%s) - T0) )); }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader > $O/gpu.txt 2>&1

# 1. parity of both forms against the oracle (each form in its own process)
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k tricubic_row_variants > $O/variants.log 2>&1
echo "variants rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log

# 2. A/B timings (device-resident, CUDA events, 10 x 10 sweeps after 3 warm-up steps)
: > $O/ab.txt
for real in double float; do
  for size in 512x256x256 1024x1024x512; do
    for rows in 1 2; do
      echo -n "rows=$rows " >> $O/ab.txt
      B200_TRICUBIC_ROWS=$rows timeout 120 bash tools/quick.sh tricubic $real $size >> $O/ab.txt 2>> $O/ab.err || echo "FAILED $real $size rows=$rows" >> $O/ab.txt
    done
  done
done
for rows in 1 2; do
  echo -n "rows=$rows " >> $O/ab.txt
  B200_TRICUBIC_ROWS=$rows timeout 150 bash tools/quick.sh tricubic double 1024x1024x1024 >> $O/ab.txt 2>> $O/ab.err || echo "FAILED double 1024^3 rows=$rows" >> $O/ab.txt
  echo -n "rows=$rows " >> $O/ab.txt
  B200_TRICUBIC_ROWS=$rows timeout 120 bash tools/quick.sh tricubic2 double 512x256x256 >> $O/ab.txt 2>> $O/ab.err || echo "FAILED tricubic2 rows=$rows" >> $O/ab.txt
done
cat $O/ab.txt
echo "ab t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log

# 3. winners per precision (sum of GLUP/s over the sizes measured)
eval $(python - <<'PY'
import re
best = {}
for line in open("gpurun_out/r1r/ab.txt"):
    m = re.match(r"rows=(\d) (tricubic2?) \S+ (double|float) niters=\d+ ([\d.]+) frac", line)
    if not m:
        continue
    rows, _, real, v = m.groups()
    best.setdefault(real, {}).setdefault(rows, 0.0)
    best[real][rows] += float(v)
for real, env in (("double", "F64"), ("float", "F32")):
    d = best.get(real, {})
    w = "2" if d.get("2", 0.0) > 1.02 * d.get("1", 0.0) else "1"
    print(f"export B200_TRICUBIC_ROWS_{env}={w};")
PY
)
echo "winners: F64=$B200_TRICUBIC_ROWS_F64 F32=$B200_TRICUBIC_ROWS_F32" | tee $O/winners.txt

# 4. one ncu --set full capture of each double form (evidence for the shared-memory / FP64 pipe argument)
for rows in 1 2; do
  B200_TRICUBIC_ROWS=$rows timeout 200 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 1 -f \
    -o $O/prof_tricubic_double_rows$rows python bench.py --test tricubic --real double --size 512x256x256 --steps 2 --warmup 3 \
    --suite none --no-e2e --no-cpu > $O/ncu_rows$rows.log 2>&1
done
if [ "$B200_TRICUBIC_ROWS_F32" = "2" ] && [ $(left) -gt 400 ]; then
  B200_TRICUBIC_ROWS=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 12 -c 1 -f \
    -o $O/prof_tricubic_float_rows2 python bench.py --test tricubic --real float --size 512x256x256 --steps 2 --warmup 3 \
    --suite none --no-e2e --no-cpu > $O/ncu_f_rows2.log 2>&1
fi
echo "ncu t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log

# 5. every GPU test that touches tricubic, under the winning setting (= the build after the defaults are flipped)
timeout 400 python -m pytest tests -q -m gpu -k "tricubic" -x > $O/pytest_tricubic.log 2>&1
echo "pytest tricubic rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -3 $O/pytest_tricubic.log

# 6. the bench line (headline + suite table) if time is left, else the headline alone
if [ $(left) -gt 240 ]; then
  timeout $(( $(left) - 20 )) python bench.py > $O/bench.json 2> $O/bench.err
else
  timeout 150 python bench.py --suite small > $O/bench.json 2> $O/bench.err
fi
echo "bench rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -c 600 $O/bench.json
