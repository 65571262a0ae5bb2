"""Multi-GPU tests (need >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests -m gpu`; skipped on
one GPU).  The correctness bar for every decomposition is bit-identity with the single-GPU run of
the same kernels -- which in turn is parity-checked against the oracle."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def ngpus():
    import torch
    return torch.cuda.device_count()


CASES = [("laplacian", (130, 20, 31)), ("wave13pt", (128, 24, 40)), ("lapgsrb", (64, 20, 33)),
         ("tricubic", (64, 18, 29)), ("tricubic2", (64, 18, 29)), ("uxx1", (66, 18, 27)),
         ("divergence", (64, 18, 27)), ("gradient", (64, 18, 27)), ("vecadd", (64, 18, 27)),
         ("jacobi", (130, 211, 1)), ("gaussblur", (128, 190, 1)), ("gameoflife", (66, 175, 1)),
         ("matvec", (128, 301, 1)), ("sincos", (32, 18, 21)), ("matmul", (130, 70, 150))]
SCAL = {"laplacian": [0.3, 0.1], "wave13pt": [0.6, -0.03, 0.09], "lapgsrb": [0.5, 0.03, 0.02, -0.01],
        "uxx1": [0.4, -0.2], "divergence": [0.6, -0.2, 0.5], "gradient": [0.6, -0.2, 0.5],
        "jacobi": [0.5, 0.1, 0.02], "gaussblur": [0.6, 0.2, 0.1, 0.05, 0.03, 0.01]}


@pytest.mark.parametrize("g", [2, 3, 4, 8])
def test_context_slabs_bitwise(pkg, oracle, g, monkeypatch):
    """Single process, g GPUs (what B200_NGPUS=g does in the C drivers): peer-pointer halo push.
    Buffers start NaN-poisoned and output arrays are loaded shell-only, as the drivers do."""
    if ngpus() < g:
        pytest.skip(f"needs {g} GPUs")
    monkeypatch.setenv("B200_POISON", "1")
    one, many = pkg.Context(1), pkg.Context(g)
    try:
        for real in ("double", "float"):
            for test, (nx, ny, ns) in CASES:
                info = pkg.test_info(test)
                ext = ns if info["ndims"] == 3 else ny
                if ext < g * (info["zghost_lo"] + info["zghost_hi"] + 1):
                    continue
                rng = np.random.default_rng(5)
                arrays = [rng.uniform(-1, 1, oracle.array_len(test, q, nx, ny, ns)).astype(
                    np.float64 if real == "double" else np.float32) for q in range(info["narrays"])]
                a = [x.copy() for x in arrays]
                b = [x.copy() for x in arrays]
                sa, _ = one.run_on_host_arrays(test, real, nx, ny, ns, SCAL.get(test, []), a, 5)
                sb, st = many.run_on_host_arrays(test, real, nx, ny, ns, SCAL.get(test, []), b, 5)
                assert sa == sb and st["ngpus"] == g
                for q in range(info["narrays"]):
                    assert np.array_equal(a[q], b[q]), f"{test}/{real} slot {q} differs on {g} GPUs"
    finally:
        one.destroy()
        many.destroy()


@pytest.mark.parametrize("halo", ["push", "nccl"])
def test_process_per_gpu_slabs_bitwise(halo):
    """One process per GPU (torchrun), CUDA-IPC peer pointers + device-side flags (push) or NCCL
    send/recv (nccl): gathered result == single-GPU result, bit for bit."""
    n = min(ngpus(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        str(ROOT / "tests" / "mp_slab_check.py"), "nccl", halo],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "MP_SLAB_CHECK OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
