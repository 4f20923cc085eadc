"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name.
    python tools/ncu_summary.py gpurun_out/launches.csv [out.md]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = name.split("(")[0].replace("spg::<unnamed>::", "").replace("void ", "")
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {n} | {us / 1e3:.3f} | {100 * us / total:.1f} % |")
    out.append(f"| **total** | {len(rows)} | {total / 1e3:.3f} | 100 % |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
