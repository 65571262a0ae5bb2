#!/usr/bin/env python
"""tools/ref_cuda_table.py [out.json] -- the GPU-vs-GPU baseline: the reference's OWN cuda target (naive one-thread-
per-point kernels, <test>/<test>.c under __CUDACC__), rebuilt unmodified for sm_100 by oracle/build_ref.sh into
oracle/_ref/cuda_bin, run on this B200 at the README size (512 256 256 10; 2D tests 512 65536 10) and parsed with
the grammar `benchmark` uses (kernel time = CUDA-event time per launch printed by __wrap_cudaLaunchKernel,
<test>/cuda/cuda_profiling.cu:214-251).  Reported beside the b200 kernels in profiles/; never a bench value of ours."""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
BIN = ROOT / "oracle" / "_ref" / "cuda_bin"
NUM = r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?"
TWO_D = {"gameoflife", "gaussblur", "matvec"}


def main():
    from pkgload import load_pkg
    pkg = load_pkg()
    out = []
    tests = sorted({p.name.rsplit("_", 1)[0] for p in BIN.glob("*_double")})
    for t in tests:
        kf = BIN / f"{t}.kernel"
        kernel = kf.read_text().strip() if kf.exists() else t
        for real in ("double", "float"):
            exe = BIN / f"{t}_{real}"
            args = ["512", "65536", "10"] if t in TWO_D else ["512", "256", "256", "10"]
            env = dict(os.environ, PROFILING_FNAME=kernel)
            row = {"test": t, "real": real, "size": "x".join(args[:-1]), "impl": "reference cuda target, sm_100"}
            try:
                p = subprocess.run([str(exe)] + args, capture_output=True, text=True, env=env, timeout=300)
                kt = [float(v) for v in re.findall(rf"{re.escape(kernel)} kernel time = ({NUM})", p.stdout)]
                ct = re.search(rf"compute time = ({NUM}) sec", p.stdout)
                fm = re.search(rf"final mean = ({NUM})", p.stdout)
                regs = re.search(rf"{re.escape(kernel)} regcount = (\d+)", p.stdout)
                if p.returncode != 0 or not kt:
                    row["error"] = f"rc={p.returncode} {p.stderr[-120:]}"
                else:
                    dims = [int(a) for a in args[:-1]] + ([1] if t in TWO_D else [])
                    lups = pkg.interior_points(t, *dims)
                    sec = sum(kt) / len(kt)
                    info = pkg.test_info(t)
                    bpl = (info["nread"] + info["nwritten"]) * (8 if real == "double" else 4)
                    row.update(us_per_sweep=round(sec * 1e6, 1), glups=round(lups / sec / 1e9, 2),
                               gbs=round(lups * bpl / sec / 1e9, 1), regs=int(regs.group(1)) if regs else None,
                               t_comp=float(ct.group(1)) if ct else None, f_mean=float(fm.group(1)) if fm else None)
            except Exception as e:      # noqa: BLE001
                row["error"] = str(e)[:160]
            print(row, flush=True)
            out.append(row)
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
