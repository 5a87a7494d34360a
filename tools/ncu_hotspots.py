"""Top stall sites of one kernel from an ncu report: `ncu -i rep --page source --csv --print-source sass` parsed and
printed as (address, sass, samples, dominant stall reasons).  usage: python tools/ncu_hotspots.py rep.ncu-rep [top_n]"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = list(csv.DictReader(lines[start:]))
    tot = sum(int(r["# Samples"] or 0) for r in rd)
    stall_cols = [c for c in rd[0] if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[c] or 0) for r in rd) for c in stall_cols}
    print("total samples", tot, "instructions", len(rd))
    print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = {id(r): i for i, r in enumerate(rd)}
    for r in sorted(rd, key=lambda r: -int(r["# Samples"] or 0))[:top]:
        n = int(r["# Samples"] or 0)
        why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("%5d %5.1f%%  #%-5d %-70s %s" % (n, 100.0 * n / max(tot, 1), idx[id(r)], r["Source"][:70], " ".join("%s=%d" % (w, k) for k, w in why if k)))

if __name__ == "__main__":
    main()
