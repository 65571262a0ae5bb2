#!/bin/bash
O=gpurun_out/${OUT:-r2s}
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 > $O/bench_n${NG:-2}.json 2> $O/bench_n${NG:-2}.err
echo "bench rc=$?"; tail -3 $O/bench_n${NG:-2}.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("$O/bench_n${NG:-2}.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"], "parity", d["parity"]["multi_eq_single"])
PY
