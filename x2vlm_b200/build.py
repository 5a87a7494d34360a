"""Build libx2k.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

The built .so lives at x2vlm_b200/lib/libx2k.so: git-ignored, but shipped to the GPU box by gpurun.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libx2k.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for dep in [path] + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ) + [os.path.join(HERE, "..", "include", "x2k.h")]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _parts(src):
    """Number of -DX2K_<NAME>_PART=i compilations a source asks for (marker: `// X2K_BUILD_PARTS: n`), else 0."""
    with open(src) as fh:
        for line in fh:
            if "X2K_BUILD_PARTS:" in line:
                return int(line.split("X2K_BUILD_PARTS:")[1].split()[0])
    return 0


def _jobs():
    jobs = []
    for src in _sources():
        n = _parts(src)
        base = os.path.basename(src)[:-3]
        if n:
            jobs += [(src, "%s_p%d" % (base, i), ["-DX2K_%s_PART=%d" % (base.upper(), i)]) for i in range(n)]
        else:
            jobs.append((src, base, []))
    return jobs


def _compile_one(nvcc, job, verbose):
    src, name, defs = job
    obj = os.path.join(OBJDIR, name + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, res.stderr if verbose else ""


def build_lib(verbose=False, force=False):
    """Compile every csrc/*.cu for sm_100a and link x2vlm_b200/lib/libx2k.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    jobs = _jobs()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, len(jobs))) as ex:
        results = list(ex.map(lambda j: _compile_one(nvcc, j, verbose), jobs))
    objs = [r[0] for r in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build_lib(verbose="-v" in sys.argv, force="-f" in sys.argv))
