#!/usr/bin/env python
"""tools/suite_table.py <bench.json> <out.md> [<ref_cuda_target.json>] -- the `suite` key of a bench.py line as the
markdown table kept under profiles/ (same layout as tools/collect_profiles.py step 5)."""
import json
import sys
from pathlib import Path

src, out = Path(sys.argv[1]), Path(sys.argv[2])
b = json.loads(src.read_text().strip().splitlines()[-1])
ref = {}
if len(sys.argv) > 3 and Path(sys.argv[3]).exists():
    for r in json.loads(Path(sys.argv[3]).read_text()):
        ref[(r["test"], r["real"])] = r.get("glups")
tab = {}
for r in b.get("suite", []):
    if r["test"] != "matmul":
        tab.setdefault(r["test"], {})[(r.get("cfg"), r["real"])] = r


def cell(r, g=True):
    if not r or "frac" not in r:
        return "-"
    return f"{r['glups']:.0f} ({r['frac']:.2f})" if g else f"{r['frac']:.2f}"


lines = ["| test | C1 d GLUP/s (frac) | C1 f GLUP/s (frac) | C2 d frac | C2 f frac | C3 d frac | gcc+OpenMP C1 d / f GLUP/s | reference cuda target C1 d / f GLUP/s |",
         "|---|---|---|---|---|---|---|---|"]
for t, v in tab.items():
    c1d, c1f = v.get(("C1", "double")), v.get(("C1", "float"))
    lines.append(f"| {t} | {cell(c1d)} | {cell(c1f)} | {cell(v.get(('C2', 'double')), False)} | {cell(v.get(('C2', 'float')), False)} | "
                 f"{cell(v.get(('C3', 'double')), False)} | {(c1d or {}).get('cpu_glups', '-')} / {(c1f or {}).get('cpu_glups', '-')} | "
                 f"{ref.get((t, 'double'), '-')} / {ref.get((t, 'float'), '-')} |")
for r in b.get("suite", []):
    if r["test"] == "matmul" and "tflops" in r:
        lines.append(f"| matmul {r['real']} 8192^3 | {r['tflops']} TFLOP/s | cuBLAS {r['cublas_tflops']} TFLOP/s | ratio {r['vs_cublas']} | | | | |")
out.write_text(
    f"Per-test roofline table of `profiles/{src.name}` (`suite` key).  C1 = 512x256x256 (2D tests 512x65536), C2 = 1024x1024x512, "
    f"C3 = 1024^3 double; frac = algorithmic bytes per sweep / time / {b['roofline']['peak']} GB/s (measured copy ceiling, "
    f"MEASURED_PEAKS.json of the run).\n\n" + "\n".join(lines) + "\n")
print(out.read_text())
