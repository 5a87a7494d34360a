"""Per-object SASS opcode census of libx2k's kernels (runs where the objects were built; no GPU needed):
tcgen05.mma -> UTCHMMA (.2CTA for cta_group::2), tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG / UTMASTG, bulk reductions /
vector atomics -> REDG, legacy tensor path -> HMMA (must stay 0).  usage: python tools/sass_census.py > profiles/rNN_sass_census.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "x2vlm_b200", "build")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "REDG", "HMMA"]
print("| object | kernels | " + " | ".join("`%s`" % o for o in OPS) + " |")
print("|---|---:|" + "---:|" * len(OPS))
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    nk = len(re.findall(r"^\s*Function : ", sass, re.M))
    counts = []
    for o in OPS:
        if o == "UTCHMMA":
            counts.append(len(re.findall(r"\bUTCHMMA(?!\.2CTA)\b", sass)))
        elif o == "HMMA":
            counts.append(len(re.findall(r"\bHMMA\b", sass)))
        else:
            counts.append(len(re.findall(re.escape(o), sass)))
    print("| `%s` | %d | %s |" % (f, nk, " | ".join(str(c) for c in counts)))
