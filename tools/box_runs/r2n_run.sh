#!/bin/bash
# tools/r2n_run.sh -- 2 GPUs, single process: kernel durations of the halo-pushing variant vs the plain one (ncu launch list)
O=gpurun_out/${OUT:-r2n}
mkdir -p $O
B=kernelgen-perf-tests_b200/drivers/bin
for t in lapgsrb laplacian tricubic gameoflife gaussblur; do
  for real in double float; do
    if [ $t = gameoflife -o $t = gaussblur -o $t = jacobi ]; then a1="512 65536 6"; a2="512 131072 6"; else a1="512 256 256 6"; a2="512 256 512 6"; fi
    B200_INIT_THREADS=16 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/n1_${t}_$real.csv $B/${t}_$real $a1 > /dev/null 2>&1
    B200_INIT_THREADS=16 B200_NGPUS=2 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/n2_${t}_$real.csv $B/${t}_$real $a2 > /dev/null 2>&1
  done
done
python - <<PY
import csv, glob, statistics
for f in sorted(glob.glob("$O/n?_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Value" in r)
    iv, ik = hdr.index("Metric Value"), hdr.index("Kernel Name")
    vals = [float(r[iv].replace(",", "")) for r in rows if r is not hdr and "stream_kernel" in r[ik]]
    unit = next(r[hdr.index("Metric Unit")] for r in rows if r is not hdr)
    if vals: print(f.split("/")[-1], "launches", len(vals), "median", round(statistics.median(vals), 1), unit, "min", round(min(vals), 1))
PY
