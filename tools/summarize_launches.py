"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
by kernel name.  usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv, re, sys
from collections import defaultdict

UNIT = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "")
    name = re.sub(r"x2k::(\(anonymous namespace\)|<unnamed>)::", "x2k::", name)
    name = re.sub(r"at::native::", "", name)
    return name[:100]


def main():
    path = sys.argv[1]
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = defaultdict(lambda: [set(), 0.0, 0.0])  # ids, ns, dram bytes
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1)
        k = short(r["Kernel Name"])
        agg[k][0].add(r["ID"])
        if r["Metric Name"] == "gpu__time_duration.sum":
            agg[k][1] += v
        elif r["Metric Name"].startswith("dram__bytes"):
            agg[k][2] += v
    total = sum(v[1] for v in agg.values())
    have_dram = any(v[2] for v in agg.values())
    out = ["| kernel | launches | total ms | share | avg us |" + (" dram MB / launch | dram GB/s |" if have_dram else ""),
           "|---|---:|---:|---:|---:|" + ("---:|---:|" if have_dram else "")]
    for k, (ids, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        n = len(ids)
        row = "| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k, n, ns / 1e6, 100 * ns / total, ns / n / 1e3)
        if have_dram:
            row += " %.2f | %.0f |" % (by / n / 1e6, by / ns if ns else 0.0)
        out.append(row)
    out.append("| **total** | %d | %.3f | 100%% | |" % (sum(len(v[0]) for v in agg.values()), total / 1e6) + (" | |" if have_dram else ""))
    fam = defaultdict(lambda: [0, 0.0, 0.0])
    for k, (ids, ns, by) in agg.items():
        f = "x2k GEMM (tcgen05)" if "gemm_tcgen05" in k else "x2k attention (tcgen05)" if "attn_" in k else \
            "x2k row / optimizer kernels" if k.startswith("x2k::") else "NCCL" if "nccl" in k.lower() else "torch glue (losses, heads, index ops)"
        fam[f][0] += len(ids); fam[f][1] += ns; fam[f][2] += by
    out += ["", "| family | launches | total ms | share |" + (" avg dram MB / launch |" if have_dram else ""), "|---|---:|---:|---:|" + ("---:|" if have_dram else "")]
    for f, (n, ns, by) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.3f | %.1f%% |" % (f, n, ns / 1e6, 100 * ns / total) + (" %.2f |" % (by / n / 1e6) if have_dram else ""))
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
