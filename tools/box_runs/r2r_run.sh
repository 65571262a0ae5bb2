#!/bin/bash
# tools/r2r_run.sh -- 8 GPUs: multi-GPU bit-identity tests (2/3/4/8 slabs), bench --gpus 8 (parity block), scaling suite N=8 and N=1
O=gpurun_out/${OUT:-r2r}
mkdir -p $O
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_drivers.py::test_multi_gpu_driver_matches_single" -x -q -m gpu > $O/pytest_multi.log 2>&1
echo "pytest multi rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log; tail -3 $O/pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
echo "bench n8 rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/scaling_suite.py $O/scaling_n8.json > /dev/null 2> $O/err_n8.txt
echo "suite n8 rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 tools/scaling_suite.py $O/scaling_n4.json --tests laplacian,wave13pt,lapgsrb,jacobi,gaussblur,gameoflife,tricubic > /dev/null 2> $O/err_n4.txt
timeout 300 python tools/scaling_suite.py $O/scaling_n1.json > /dev/null 2> $O/err_n1.txt
echo "suite n1 rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
python - <<PY
import json
def load(f):
    d=json.load(open(f)); rows = d["rows"] if isinstance(d,dict) else d
    return {(r["test"],r["real"]):r for r in rows}
a=load("$O/scaling_n1.json"); b=load("$O/scaling_n8.json"); c=load("$O/scaling_n4.json")
for k in a:
    print(f"{k[0]:11s} {k[1]:6s} N1 {a[k]['glups']:7.1f} N4 {c[k]['glups'] if k in c else 0:8.1f} N8 {b[k]['glups']:8.1f} eff8 {b[k]['glups']/(8*a[k]['glups']):.3f}")
d = json.loads(open("$O/bench_n8.json").read().strip().splitlines()[-1])
print("bench n8", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"]["value"], "parity", d["parity"]["multi_eq_single"], d["parity"]["bytes_compared"])
PY
