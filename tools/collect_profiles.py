#!/usr/bin/env python
"""Turn the raw artefacts of tools/make_profiles.sh (gpurun_out/*_<tag>.*) into the tracked evidence
under profiles/: launch list + per-kernel shares, steady-state DRAM traffic per launch of the
headline kernel (-> profiles/traffic.json, read by bench.py), the ncu `--set full` key metrics."""
import csv
import json
import shutil
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)


def rows(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


# 1. launch list
shutil.copy(G / f"launches_{tag}.csv", P / f"{tag}_launches.csv")
per = defaultdict(list)
for r in rows(G / f"launches_{tag}.csv"):
    if r["Metric Name"] == "gpu__time_duration.sum":
        per[r["Kernel Name"]].append(float(r["Metric Value"]))
tot = sum(sum(v) for v in per.values())
with open(P / f"{tag}_launch_summary.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 400 python bench.py --steps 20 --warmup 3 --suite none --no-cpu\n")
    f.write("# (serialised launches: compare SHARES; the e2e leg's memcpys are not kernels)\n")
    f.write(f"{'kernel':90s} {'launches':>8s} {'mean_us':>9s} {'share':>7s}\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k[:90]:90s} {len(v):8d} {sum(v) / len(v) / 1e3:9.2f} {sum(v) / tot:7.3f}\n")

# 2. steady-state traffic of the headline kernel
by = defaultdict(dict)
for r in rows(G / f"traffic_{tag}.csv"):
    by[r["ID"]][r["Metric Name"]] = float(r["Metric Value"])
    by[r["ID"]]["name"] = r["Kernel Name"]
rd = [m["dram__bytes_read.sum"] for m in by.values()]
wr = [m["dram__bytes_write.sum"] for m in by.values()]
du = [m["gpu__time_duration.sum"] for m in by.values()]
traffic = (sum(rd) + sum(wr)) / len(rd)
with open(P / f"{tag}_traffic.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none\n")
    f.write("#     -k regex:stream_kernel -s 30 -c 20 python bench.py --steps 5 --warmup 3 --suite none --no-e2e --no-cpu\n")
    f.write(f"kernel: {next(iter(by.values()))['name']}\n")
    f.write(f"launches {len(rd)}  mean duration {sum(du) / len(du) / 1e3:.2f} us\n")
    f.write(f"dram read  per launch {sum(rd) / len(rd) / 1e6:.1f} MB\n")
    f.write(f"dram write per launch {sum(wr) / len(wr) / 1e6:.1f} MB\n")
    f.write(f"dram total per launch {traffic / 1e6:.1f} MB   (algorithmic: 32,260,032 LUP x 24 B = 774.2 MB)\n")
tj = P / "traffic.json"
d = json.loads(tj.read_text()) if tj.exists() else {}
d["wave13pt_double_512x256x256"] = traffic
d["_source"] = f"profiles/{tag}_traffic.txt (steady state, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
tj.write_text(json.dumps(d, indent=1) + "\n")

# 3. full-set capture: key metrics
out = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(G / f"prof_wave13pt_{tag}.ncu-rep")],
                     capture_output=True, text=True).stdout
det = subprocess.run(["ncu", "-i", str(G / f"prof_wave13pt_{tag}.ncu-rep"), "--page", "details"], capture_output=True, text=True).stdout
keep = [l for l in det.splitlines() if any(k in l for k in (
    "Duration", "DRAM Throughput", "Memory Throughput", "L2 Hit Rate", "Registers Per Thread", "Dynamic Shared Memory",
    "Grid Size", "Block Size", "Executed Ipc", "Issue Slots Busy", "Achieved Occupancy", "SM Frequency", "DRAM Frequency",
    "L1/TEX Hit Rate", "Shared Memory Configuration", "Theoretical Occupancy", "Eligible Warps", "FP64"))]
(P / f"{tag}_ncu_wave13pt_double.txt").write_text(
    "# ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 30 -c 1 python bench.py --steps 3 --warmup 3 --suite none --no-e2e --no-cpu\n"
    "# NOTE: ncu flushes the caches before the launch, so dram write bytes miss the part of the output still dirty in L2\n"
    "# at kernel end; the steady-state traffic is in the _traffic.txt file.\n" + out + "\n".join(keep) + "\n")
shutil.copy(G / f"bench_{tag}.json", P / f"{tag}_bench.json")

# 4. every other kernel captured by make_profiles.sh: one line of key metrics each
reps = sorted(p for p in G.glob(f"prof_*_{tag}.ncu-rep") if "wave13pt" not in p.name)
if reps:
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py")] + [str(p) for p in reps],
                         capture_output=True, text=True).stdout
    (P / f"{tag}_ncu_all_kernels.txt").write_text(
        "# ncu --set full --clock-control none, one launch per kernel (tools/ncu_one.sh; caches flushed before the launch, so\n"
        "# dur_us is the isolated kernel and wr_MB misses what is still dirty in L2).  Columns: tools/ncu_summary.py KEYS.\n" + out)
if (G / f"ref_cuda_{tag}.json").exists():
    shutil.copy(G / f"ref_cuda_{tag}.json", P / f"{tag}_ref_cuda_target.json")

# 5. the per-test table of the bench line, as markdown
b = json.loads((G / f"bench_{tag}.json").read_text().strip().splitlines()[-1])
tab, cpu = {}, {}
for r in b.get("suite", []):
    if r["test"] == "matmul":
        continue
    tab.setdefault(r["test"], {})[(r.get("cfg"), r["real"])] = r
ref = {}
if (P / f"{tag}_ref_cuda_target.json").exists():
    for r in json.loads((P / f"{tag}_ref_cuda_target.json").read_text()):
        ref[(r["test"], r["real"])] = r.get("glups")
lines = ["| test | C1 d GLUP/s (frac) | C1 f GLUP/s (frac) | C2 d frac | C2 f frac | C3 d frac | gcc+OpenMP C1 d / f GLUP/s | reference cuda target C1 d / f GLUP/s |",
         "|---|---|---|---|---|---|---|---|"]
def cell(r, g=True):
    if not r or "frac" not in r:
        return "-"
    return f"{r['glups']:.0f} ({r['frac']:.2f})" if g else f"{r['frac']:.2f}"
for t, v in tab.items():
    c1d, c1f = v.get(("C1", "double")), v.get(("C1", "float"))
    lines.append(f"| {t} | {cell(c1d)} | {cell(c1f)} | {cell(v.get(('C2', 'double')), False)} | {cell(v.get(('C2', 'float')), False)} | "
                 f"{cell(v.get(('C3', 'double')), False)} | {(c1d or {}).get('cpu_glups', '-')} / {(c1f or {}).get('cpu_glups', '-')} | "
                 f"{ref.get((t, 'double'), '-')} / {ref.get((t, 'float'), '-')} |")
for r in b.get("suite", []):
    if r["test"] == "matmul" and "tflops" in r:
        lines.append(f"| matmul {r['real']} 8192^3 | {r['tflops']} TFLOP/s | cuBLAS {r['cublas_tflops']} TFLOP/s | ratio {r['vs_cublas']} | | | | |")
(P / f"{tag}_suite_table.md").write_text(
    f"Per-test roofline table of `profiles/{tag}_bench.json` (`suite` key).  C1 = 512x256x256 (2D tests 512x65536), C2 = 1024x1024x512, "
    f"C3 = 1024^3 double; frac = algorithmic bytes per sweep / time / {b['roofline']['peak']} GB/s (measured copy ceiling).\n\n" + "\n".join(lines) + "\n")
print((P / f"{tag}_suite_table.md").read_text())
print(open(P / f"{tag}_launch_summary.txt").read())
print(open(P / f"{tag}_traffic.txt").read())
