"""Phase timelines (clock64 stamps of one CTA) of the attention backward kernels — developer tool.
Build the traced library here:   python tools/trace_attn.py --build
run on the GPU:                  X2K_LIB=x2vlm_b200/lib/libx2k_trace.so python tools/trace_attn.py"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--build" in sys.argv:
    from x2vlm_b200 import build as b
    b.build_lib()
    traced = {"attn_pack": "-DX2K_PACK_TRACE", "attn": "-DX2K_ATTN_TRACE"}
    objs = [os.path.join(b.OBJDIR, f) for f in os.listdir(b.OBJDIR)
            if f.endswith(".o") and f[:-2] not in traced and not f.endswith("_trace.o")]
    flags = [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    tobjs = []
    for base, d in traced.items():
        tobj = os.path.join(b.OBJDIR, base + "_trace.o")
        subprocess.check_call([b._nvcc()] + flags + [d, "-c", os.path.join(b.CSRC, base + ".cu"), "-o", tobj])
        tobjs.append(tobj)
    subprocess.check_call([b._nvcc(), "-shared", "-o", os.path.join(b.LIBDIR, "libx2k_trace.so")] + objs + tobjs + ["-lcudart"])
    for t in tobjs:
        os.remove(t)
    print("built libx2k_trace.so"); sys.exit(0)
import runpy
import torch
from x2vlm_b200 import _capi as C


def dump(fn, names, threads=(("thread 0", 0), ("thread 200", 64))):
    buf = (ctypes.c_longlong * 192)()
    f = getattr(C.lib(), fn)
    f.argtypes = [ctypes.c_void_p]
    assert f(buf) == 0
    for who, base in threads:
        t = [buf[base + i] for i in range(64)]
        print(" ", who)
        if t[0] == 0:
            t[0] = min(x for x in t if x > 0)
        prev = t[0]
        for i in sorted(names):
            if t[i] == 0 or t[i] < t[0]:
                continue
            print("    %-45s +%7d cyc  (delta %6d)" % (names[i], t[i] - t[0], t[i] - prev))
            prev = t[i]


def pack_names():
    names = {0: "start", 1: "prologue done (tmem alloc, barriers)", 62: "dK/dV drained", 63: "end"}
    for ci in range(6):
        for k, n in enumerate(["item start", "row stats done", "Q/dO (+K/V) landed", "S ready", "pass 1 done", "dP ready", "pass 2 done",
                               "sync before final MMAs", "dQ/dK/dV MMAs done", "dQ drained + sync"]):
            names[2 + ci * 10 + k] = "item %d: %s" % (ci, n)
    return names


def attn_names():
    names = {0: "start", 1: "prologue done (tmem alloc, barriers)", 2: "lse / delta loaded", 60: "dQ drained", 61: "end"}
    for t in range(4):
        for k, n in enumerate(["operands landed, S/dP MMAs go", "S/dP ready", "P/dS pass done", "synced", "dQ/dK/dV MMAs issued"]):
            names[3 + t * 6 + k] = "tile kb%d qb%d: %s" % (t >> 1, t & 1, n)
    for kb in range(2):
        for k, n in enumerate(["dK/dV ready", "dK/dV drained", "synced"]):
            names[40 + kb * 3 + k] = "key block %d: %s" % (kb, n)
    for i in range(2):
        for k, n in enumerate(["next S/dP issued", "P/dS stored (barrier passed)", "dQ/dK/dV MMAs issued"]):
            names[46 + i * 4 + k] = "MMA warp, last %s tile: %s" % ("even" if i == 0 else "odd", n)
    return names


def fwd_names():
    return dict(enumerate(["start", "prologue done, S MMAs issued (thread 0)", "S ready", "pass 1 done (bias, max)", "synced", "pass 2 done (exp, P)",
                           "sums exchanged, P·V issued", "O ready", "O stored", "end"]))


only = os.environ.get("X2K_ATTN_CASE")
for case, fn, names in (("beit", "x2k_debug_attn_trace", attn_names()), ("fus-self", "x2k_debug_pack_trace", pack_names()),
                        ("cross", "x2k_debug_pack_trace", pack_names())):
    if only and only != case:
        continue
    os.environ["X2K_ATTN_CASE"] = case
    sys.argv = [sys.argv[0]]
    runpy.run_path(os.path.join(ROOT, "tools", "profile_attn.py"))
    torch.cuda.synchronize()
    print("TRACE", case)
    if case == "beit":
        print(" forward (attn_fwd_kernel)")
        dump("x2k_debug_attn_fwd_trace", fwd_names())
        print(" backward (attn_bwd_kernel)")
    if case == "beit":
        dump(fn, names, (("thread 0", 0), ("thread 200", 64), ("thread 256 (MMA warp)", 128)))
    else:
        dump(fn, names)
