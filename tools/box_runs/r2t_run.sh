#!/bin/bash
O=gpurun_out/r2t
mkdir -p $O
cat > /tmp/mm_one.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from pkgload import load_pkg
pkg = load_pkg(); pkg.load()
n = int(sys.argv[1]); real = sys.argv[2]
dt = torch.float32 if real == "float" else torch.float64
A = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1; B = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1; C = torch.zeros(n * n, device="cuda", dtype=dt)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    pkg.capi.sweep_loop("matmul", real, n, n, n, [], [A.data_ptr(), B.data_ptr(), C.data_ptr()], 1, stream=s)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:matmul_tc05_kernel -s 1 -c 1 -f -o $O/prof_matmul_tc05_4096 python /tmp/mm_one.py 4096 float > $O/ncu.log 2>&1
echo "ncu rc=$?"
ls -la $O
