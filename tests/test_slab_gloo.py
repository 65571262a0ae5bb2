"""CPU test (world_size 2 and 3, gloo) of the N>1 host logic: slab partition, ghost depths,
which planes are exchanged and in which buffer, rotation across ranks.  The per-slab sweep is the
oracle (this is a test), so the gathered result must be bit-identical to the undivided grid."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_gloo(world):
    port = 29500 + world
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        str(ROOT / "tests" / "mp_slab_check.py"), "gloo"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MP_SLAB_CHECK OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_layout_arithmetic(pkg):
    from kernelgen_perf_tests_b200.slab import SlabLayout
    info = pkg.test_info("wave13pt")
    L = [SlabLayout(info, 64, 4, r) for r in range(4)]
    assert [(l.own_lo, l.own_hi) for l in L] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert (L[0].mem_lo, L[0].mem_hi) == (0, 18) and (L[1].mem_lo, L[1].mem_hi) == (14, 34)
    assert L[0].out_range() == (2, 16) and L[3].out_range() == (2, 16)
    assert L[1].send_lo() == (2, 4) and L[1].send_hi() == (16, 18) and L[1].recv_lo() == (0, 2)
    info = pkg.test_info("tricubic")      # asymmetric ghosts: 1 below, 2 above
    a, b = SlabLayout(info, 40, 2, 0), SlabLayout(info, 40, 2, 1)
    assert (a.mem_lo, a.mem_hi, b.mem_lo, b.mem_hi) == (0, 22, 19, 40)
    assert a.send_hi_cnt == 1 and b.send_lo_cnt == 2
    with pytest.raises(ValueError):
        SlabLayout(info, 6, 4, 1)
