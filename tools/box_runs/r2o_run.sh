#!/bin/bash
# 2 GPUs: multi-GPU bit-identity tests, then kernel durations of PUSH vs plain (ncu launch list), then scaling suite N=2
O=gpurun_out/${OUT:-r2o}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_drivers.py::test_multi_gpu_driver_matches_single" -x -q -m gpu > $O/pytest_multi.log 2>&1
echo "pytest multi rc=$?"; tail -4 $O/pytest_multi.log
[ -n "$DUR" ] && OUT=${OUT:-r2o}/durations bash tools/box_runs/r2n_run.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/scaling_suite.py $O/scaling_n2.json --tests laplacian,wave13pt,lapgsrb,jacobi,gaussblur,gameoflife,tricubic > /dev/null 2> $O/err_n2.txt
python tools/scaling_suite.py $O/scaling_n1.json --tests laplacian,wave13pt,lapgsrb,jacobi,gaussblur,gameoflife,tricubic > /dev/null 2> $O/err_n1.txt
python - <<PY
import json
def load(f):
    d=json.load(open(f)); rows = d["rows"] if isinstance(d,dict) else d
    return {(r["test"],r["real"]):r for r in rows}
a=load("$O/scaling_n1.json"); b=load("$O/scaling_n2.json")
for k in a:
    print(f"{k[0]:11s} {k[1]:6s} N1 {a[k]['glups']:7.1f} N2 {b[k]['glups']:7.1f} eff {b[k]['glups']/(2*a[k]['glups']):.3f}")
PY
