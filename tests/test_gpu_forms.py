"""GPU parity of every alternative kernel FORM the product library can choose at launch time (tile policy): each form
runs in its own process (the library reads such switches once), against the CPU oracle, and all forms of one stencil must
produce byte-identical outputs."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from parity_util import TOL

pytestmark = [pytest.mark.gpu]
HERE = Path(__file__).resolve().parent


def _run(tests, env):
    p = subprocess.run([sys.executable, str(HERE / "op_variant_check.py")] + tests, capture_output=True, text=True,
                       env=dict(os.environ, **env), timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_small_tile_forms_float():
    """Half-height float tiles of laplacian / wave13pt / divergence / gradient / lapgsrb (B200_TILE_POLICY=2 forces
    them, =1 lets the decomposition model choose): same per-point arithmetic as the default forms, hence the same bytes."""
    tests = ["laplacian", "wave13pt", "divergence", "gradient", "lapgsrb"]
    base = _run(tests, {"B200_TILE_POLICY": "0"})
    for pol in ("2", "1"):
        out = _run(tests, {"B200_TILE_POLICY": pol})
        for real in ("float", "double"):
            assert out["worst"][real] <= TOL[real], out
        assert out["sha"] == base["sha"], f"policy {pol}: outputs differ from the default forms"
