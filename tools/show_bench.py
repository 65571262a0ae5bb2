#!/usr/bin/env python
"""Compact view of a bench.py JSON line: headline + per-test roofline fractions (d/f, C1 | C2)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s = d.pop("suite", [])
e = d.get("e2e") or {}
print(f"value {d['value']:.1f} {d['unit']}  frac {d['roofline']['frac']:.3f}  ms/step {d['ms_per_step']:.3f}  "
      f"e2e {e.get('value', 0):.2f}  launches {d['gpu_launches']}  clocks {d.get('clocks')}")
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
tab = {}
for r in s:
    if r["test"] == "matmul":
        print("matmul", r)
        continue
    tab.setdefault(r["test"], {})[(r["cfg"], r["real"])] = r.get("frac", r.get("error", "?"))
for t, v in tab.items():
    print(f"{t:11s} C1 d {v.get(('C1','double'))!s:7} f {v.get(('C1','float'))!s:7} | C2 d {v.get(('C2','double'))!s:7} f {v.get(('C2','float'))!s:7}"
          f" | C3 d {v.get(('C3','double'))!s:7}")
