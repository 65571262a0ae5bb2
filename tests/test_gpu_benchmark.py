"""The reference's own `benchmark` script driving the b200 target ON THE GPU (missing #1):

    ./benchmark 512 256 256 10 2 b200 gcc        (benchmark:12-24; run_test -> `make test`, benchmark:297)

in a writable copy of the suite (staged by oracle/build_ref.sh under oracle/_ref/suite -- git-ignored, travels to the
GPU box like the other _ref artefacts) with the b200 overlay installed (suite_overlay/install_overlay.sh).  Checked on
the AVERAGE table and the report blocks the script prints:
  * no FAIL / N/A cell in any b200 row of the 15 tests the library implements,
  * i_mean equal (as printed) and f_mean equal (to the digits printed) between the b200 and gcc rows of the 12 C tests,
  * t_krn / nreg_krn numeric for b200 (the cuda-like branch of report_test, benchmark:174-178,210-216),
  * "Times for chart >>" carries gcc/b200 ratios, `mkchart` turns the report into its bargraph input, `mktable` into its
    LaTeX rows (mkchart:120-169, mktable:41-76),
  * the appended GLUP/s and %roof columns are numeric for b200.
The table is written to gpurun_out/benchmark_b200_gcc.txt (copied to profiles/ by the builder)."""
import os
import re
import shutil
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
SUITE = ROOT / "oracle" / "_ref" / "suite"
B200_TESTS = ["laplacian", "wave13pt", "divergence", "gradient", "uxx1", "lapgsrb", "jacobi", "gaussblur",
              "gameoflife", "tricubic", "tricubic2", "vecadd", "matvec", "sincos", "matmul"]
C_TESTS = [t for t in B200_TESTS if t not in ("jacobi", "sincos", "matmul")]      # gcc target builds without gfortran
COLS = ["test", "target", "i_mean", "t_init", "t_alloc", "t_load", "t_comp", "t_krn", "nreg_krn", "t_save", "t_free",
        "f_mean", "GLUP/s", "%roof"]


def parse_tables(out: str):
    """{section: {(test, target): {column: cell}}}; sections are 'RUN #k' and 'AVERAGE'."""
    tabs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\|\s*RUN #(\d+)", line)
        if m:
            cur = tabs.setdefault(f"RUN #{m.group(1)}", {})
            continue
        if re.match(r"\|\s*AVERAGE OF", line):
            cur = tabs.setdefault("AVERAGE", {})
            continue
        cells = [c.strip() for c in line.strip().strip("|").split("|")]
        if cur is not None and len(cells) == len(COLS) and cells[1] in ("b200", "gcc"):
            cur[(cells[0], cells[1])] = dict(zip(COLS, cells))
    return tabs


@pytest.mark.skipif(not (SUITE / "benchmark").exists(), reason="oracle/_ref/suite not staged (make -C oracle ref in the build container)")
def test_real_benchmark_b200_vs_gcc(tmp_path, pkg):
    pkg.load()
    size = os.environ.get("B200_BENCHMARK_SIZE", "512 256 256 10 2").split()
    suite = tmp_path / "suite"
    shutil.copytree(SUITE, suite, symlinks=True)
    inst = ROOT / "kernelgen-perf-tests_b200" / "suite_overlay" / "install_overlay.sh"
    subprocess.run(["sh", str(inst), str(suite), str(ROOT)], check=True, capture_output=True)
    jobs = str(os.cpu_count() or 4)
    b = subprocess.run(["make", "-s", "-j", jobs, "b200"] + [f"{t}.gcc" for t in C_TESTS], cwd=suite, capture_output=True, text=True)
    assert b.returncode == 0, b.stdout[-2000:] + b.stderr[-2000:]
    env = dict(os.environ, B200_INIT_THREADS="8")          # same rand() draw order, bit-identical arrays (drivers/kg_rand.h)
    p = subprocess.run(["./benchmark"] + size + ["b200", "gcc"], cwd=suite, capture_output=True, text=True, env=env, timeout=2400)
    out = p.stdout
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    (outdir / "benchmark_b200_gcc.txt").write_text(out + "\n---- stderr ----\n" + p.stderr[-4000:])
    assert p.returncode == 0, p.stderr[-2000:]
    tabs = parse_tables(out)
    avg = tabs["AVERAGE"]
    num = re.compile(r"^[-+]?[0-9]*\.?[0-9]+([eE][-+]?[0-9]+)?$")
    for t in B200_TESTS:
        assert f"Found test {t}" in out
        row = avg[(t[:10], "b200")]
        for c in COLS[2:]:
            if t == "matmul" and c in ("GLUP/s", "%roof"):
                continue                      # tensor-core bound: the driver reports TFLOP/s, not GLUP/s against HBM
            assert row[c] not in ("FAIL", "N/A") and num.match(row[c]), f"{t}/b200: column {c} = {row[c]!r}\n{out[-3000:]}"
        assert float(row["t_krn"]) > 0 and float(row["nreg_krn"]) >= 16
        if t != "matmul":
            assert float(row["GLUP/s"]) > 0 and 0 < float(row["%roof"]) < 130
    for t in C_TESTS:
        g, b2 = avg[(t[:10], "gcc")], avg[(t[:10], "b200")]
        assert g["t_comp"] not in ("FAIL", "N/A")
        assert g["i_mean"] == b2["i_mean"], f"{t}: i_mean {b2['i_mean']} (b200) != {g['i_mean']} (gcc)"
        fg, fb = float(g["f_mean"]), float(b2["f_mean"])
        assert abs(fg - fb) <= 2e-6 * max(abs(fg), 1e-6) + 1e-12 or g["f_mean"] == b2["f_mean"], \
            f"{t}: f_mean {b2['f_mean']} (b200) vs {g['f_mean']} (gcc)"
    # report blocks: chart data = t_krn(b200) t_krn(gcc) => ratio, for every test both targets ran
    m = re.search(r"Times for chart >>\n(.*)<< Times for chart", out, re.S)
    assert m, out[-2000:]
    chart = {ln.split()[0]: ln.split() for ln in m.group(1).strip().splitlines()}
    for t in C_TESTS:
        w = chart[t]
        assert w[3] == "=>" and float(w[1]) > 0 and float(w[2]) > 0 and abs(float(w[4]) - float(w[2]) / float(w[1])) <= 1e-6 * float(w[4])
        assert float(w[4]) > 1.0, f"{t}: gcc/b200 kernel-time ratio {w[4]}"
    assert "Data for table >>" in out and "\\multicolumn{2}{l|}{b200}" in out
    report = suite / "report.b200"
    report.write_text(out)
    tab = subprocess.run(["perl", "./mktable", str(report)], cwd=suite, capture_output=True, text=True)
    assert tab.returncode == 0 and "Error parsing" not in tab.stdout
    for t in B200_TESTS:
        assert re.search(rf"^{t} & [0-9.]+ & \d+", tab.stdout, re.M), tab.stdout
    ch = subprocess.run(["perl", "./mkchart", "-xlabel", "test", "-ylabel", "gcc / b200", "-o", "chart.pdf", str(report)], cwd=suite,
                        capture_output=True, text=True)
    assert "Error parsing" not in ch.stdout and "Cannot parse" not in ch.stdout
    assert "=cluster b200" in ch.stdout
    for t in C_TESTS:
        assert re.search(rf"^{t} [0-9.eE+-]+ [0-9.eE+-]+ => [0-9.eE+-]+", ch.stdout, re.M), ch.stdout
    (outdir / "benchmark_b200_gcc.txt").write_text(out + "\n---- mktable ----\n" + tab.stdout + "\n---- mkchart (bargraph input) ----\n" + ch.stdout)
