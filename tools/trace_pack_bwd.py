"""Phase timeline of one CTA of the packed attention backward (developer tool).
Build the traced library first:  python tools/trace_pack_bwd.py --build ; then run on the GPU with
X2K_LIB=x2vlm_b200/lib/libx2k_trace.so python tools/trace_pack_bwd.py"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--build" in sys.argv:
    from x2vlm_b200 import build as b
    b.build_lib()
    objs = [os.path.join(b.OBJDIR, f) for f in os.listdir(b.OBJDIR) if f.endswith(".o") and not f.startswith("attn_pack")]
    tobj = os.path.join(b.OBJDIR, "attn_pack_trace.o")
    subprocess.check_call([b._nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DX2K_PACK_TRACE", "-c", os.path.join(b.CSRC, "attn_pack.cu"), "-o", tobj])
    subprocess.check_call([b._nvcc(), "-shared", "-o", os.path.join(b.LIBDIR, "libx2k_trace.so")] + objs + [tobj, "-lcudart"])
    os.remove(tobj)
    print("built libx2k_trace.so"); sys.exit(0)
import torch
from x2vlm_b200 import _capi as C
import runpy
sys.argv = [sys.argv[0]]
runpy.run_path(os.path.join(ROOT, "tools", "profile_attn.py"))   # the last bwd launch is the grouped cross case
buf = (ctypes.c_longlong * 128)()
C.lib().x2k_debug_pack_trace.argtypes = [ctypes.c_void_p]
assert C.lib().x2k_debug_pack_trace(buf) == 0
for who, base in (("thread 0", 0), ("thread 200", 64)):
    t = [buf[base + i] for i in range(64)]
    t0 = t[0]
    names = {0: "start", 1: "prologue done (tmem alloc, barriers)", 62: "dK/dV drained", 63: "end"}
    for ci in range(6):
        for k, n in enumerate(["item start", "row stats done", "Q/dO (+K/V) landed", "S ready", "pass 1 done", "dP ready", "pass 2 done",
                               "sync before final MMAs", "dQ/dK/dV MMAs done", "dQ drained + sync"]):
            names[2 + ci * 10 + k] = "item %d: %s" % (ci, n)
    print(who)
    prev = t0
    for i in sorted(names):
        if t[i] >= t0 and t[i] != 0 and (i < 2 or t[i] > 0):
            if i >= 2 and i < 62 and t[i] < t[1]:
                continue
            print("  %-45s +%7d cyc  (delta %6d)" % (names[i], t[i] - t0, t[i] - prev))
            prev = t[i]
