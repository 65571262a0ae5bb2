import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from pkgload import load_pkg
pkg = load_pkg(); pkg.load()
stream = torch.cuda.current_stream().cuda_stream
def mm(A, B, m, k, n):
    C = torch.zeros(m * n, device="cuda")
    pkg.capi.sweep_loop("matmul", "float", m, k, n, [], [A.contiguous().data_ptr(), B.contiguous().data_ptr(), C.data_ptr()], 1, stream=stream)
    torch.cuda.synchronize()
    return C.view(n, m).t()          # C[m, n]
import os
dbg = int(os.environ.get("B200_TC05_DBG", "0"))
print("dbg mode", dbg)
m, k, n = 128, 32, 128
ones_A = torch.ones(m * k, device="cuda"); ones_B = torch.ones(k * n, device="cuda")
C = mm(ones_A, ones_B, m, k, n)
print("ones x ones: expect 32 everywhere; got min/max", float(C.min()), float(C.max()), "C[0,:4]", C[0, :4].tolist(), "C[:4,0]", C[:4, 0].tolist())
# A[m,k] = m  (column-major storage: A.view(k, m)[kk, mm])
A = torch.arange(m, device="cuda", dtype=torch.float32).repeat(k)        # index kk*m + mm -> mm
C = mm(A, ones_B, m, k, n)
print("A=m, B=1: expect C[m,n] = 32 m; C[:6,0]", C[:6, 0].tolist(), "C[33,5]", float(C[33, 5]), "C[127,127]", float(C[127, 127]))
# B[k,n] = n  (column-major k x n: index nn*k + kk -> nn)
B = torch.arange(n, device="cuda", dtype=torch.float32).repeat_interleave(k)
C = mm(ones_A, B, m, k, n)
print("A=1, B=n: expect C[m,n] = 32 n; C[0,:6]", C[0, :6].tolist(), "C[5,33]", float(C[5, 33]))
# A[m,k] = k, B[k,n] = (k == 3)  -> C = 3
A = torch.arange(k, device="cuda", dtype=torch.float32).repeat_interleave(m)
B = (torch.arange(k, device="cuda") == 3).float().repeat(n)
C = mm(A, B, m, k, n)
print("A=k, B=delta(k,3): expect 3; got min/max", float(C.min()), float(C.max()))

if dbg in (4, 5):
    A = (torch.arange(m * k, device="cuda") % 1000).float()      # A flat index pattern
    B = (torch.arange(k * n, device="cuda") % 1000).float() + 0.5
    C = mm(A, B, m, k, n)
    flat = C.t().contiguous().view(-1)                             # flat[(col)*128 + row] = smem word (col*128 + row)
    print("raw smem words 0..40:", flat[:40].tolist())
    print("words 128..140:", flat[128:140].tolist())
    print("words 1024..1036:", flat[1024:1036].tolist())
