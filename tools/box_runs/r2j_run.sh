#!/bin/bash
# tools/r2j_run.sh -- 2 GPUs: multi-GPU bit-identity tests, bench --gpus 2 (parity block), reference arm under torchrun
O=gpurun_out/r2j
mkdir -p $O
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_drivers.py::test_multi_gpu_driver_matches_single" "tests/test_gpu_parity.py::test_async_refused_on_multi_gpu_context" -x -q -m gpu > $O/pytest_multi.log 2>&1
echo "pytest multi rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
tail -5 $O/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench n2 rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
echo "ref n2 rc=$? t=$(( $(date +%s) - T0 ))" | tee -a $O/steps.log
timeout 300 python bench.py --suite none > $O/bench_n1.json 2> $O/bench_n1.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2", "bench_ref_n2"):
    try:
        d = json.loads(open(f"gpurun_out/r2j/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", (d.get("e2e") or {}).get("value"), "parity", d.get("parity"), "cores", (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $O/bench_n2.err
