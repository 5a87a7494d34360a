"""TEST / BENCH INFRASTRUCTURE — build the travelling copy of the reference: oracle/_ref/.

The reference (zengyan-97/X2-VLM) is a pure-Python tree with no packaging, and /root/reference only exists in the
build container.  To time and test the UNMODIFIED reference on the GPU box (bench.py --impl reference, the
`gpu_eager_baseline` leg, tests/test_gpu_reference.py) this recipe COMPILES every reference module from the sources
where they lie under /root/reference into sourceless byte-code:

    /root/reference/models/xvlm.py  ->  oracle/_ref/models/xvlm.pyc      (py_compile, same interpreter as the GPU box)

No reference source text enters the repository or the snapshot: oracle/_ref/ holds only compiler outputs, is listed
in .gitignore (stays out of history) and is NOT gpurun-ignored (travels like our own built .so files).  The two tiny
JSON vision configs the reference reads at construction time (three numbers each) are re-synthesised by
oracle/ref_shim.workdir(), not copied.  Nothing under x2vlm_b200/ imports this.

    python -m oracle.build_ref           # (re)build when /root/reference is present; no-op otherwise
"""
import hashlib
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("X2VLM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
# packages / modules of the reference that the hot path's callers import (models -> utils, dataset -> vqaTools, refTools)
TOP = ("models", "utils", "dataset", "accelerators", "vqaTools", "refTools", "optim.py", "scheduler.py", "Pretrain.py",
       "Retrieval.py")


def _sources():
    for top in TOP:
        p = os.path.join(SRC_ROOT, top)
        if os.path.isfile(p):
            yield p
        elif os.path.isdir(p):
            for d, _dirs, files in os.walk(p):
                for f in sorted(files):
                    if f.endswith(".py"):
                        yield os.path.join(d, f)


def build(verbose=False):
    """Compile the reference into oracle/_ref (idempotent; stamp = sha1 of the sources + interpreter magic).
    Returns the output directory, or None when /root/reference is absent (GPU box: the prebuilt copy is used)."""
    if not os.path.isdir(os.path.join(SRC_ROOT, "models")):
        return None
    srcs = list(_sources())
    h = hashlib.sha1(sys.version.encode())
    for s in srcs:
        h.update(s.encode())
        with open(s, "rb") as fh:
            h.update(fh.read())
    stamp = os.path.join(OUT, ".stamp")
    if os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return OUT
    n = 0
    for s in srcs:
        rel = os.path.relpath(s, SRC_ROOT)
        dst = os.path.join(OUT, rel + "c")  # x.py -> x.pyc next to where x.py would be: the sourceless-import layout
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path recorded in the code object / tracebacks — keep the reference-relative name
        py_compile.compile(s, cfile=dst, dfile=os.path.join("reference", rel), doraise=True, optimize=0)
        n += 1
    with open(stamp, "w") as fh:
        fh.write(h.hexdigest())
    if verbose:
        print("oracle/_ref: %d modules compiled from %s" % (n, SRC_ROOT))
    return OUT


if __name__ == "__main__":
    out = build(verbose=True)
    print(out if out else "reference tree not present at %s: nothing built" % SRC_ROOT)
