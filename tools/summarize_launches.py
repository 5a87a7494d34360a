"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv, re, sys
from collections import defaultdict

def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "")
    name = re.sub(r"x2k::\(anonymous namespace\)::", "x2k::", name)
    name = re.sub(r"at::native::", "", name)
    return name[:110]

def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        k = short(r["Kernel Name"])
        agg[k][0] += 1; agg[k][1] += ns; total += ns
    out = ["| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k, n, ns / 1e6, 100 * ns / total, ns / n / 1e3))
    out.append("| **total** | %d | %.3f | 100%% | |" % (sum(v[0] for v in agg.values()), total / 1e6))
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)

if __name__ == "__main__":
    main()
