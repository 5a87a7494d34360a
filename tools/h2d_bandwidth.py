"""Pinned host<->device copy bandwidth and host fp32->bf16 conversion time at the step's batch size (developer tool)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (8, 54, 256):
    n = mb * 1024 * 1024 // 4
    h = torch.randn(n).pin_memory()
    d = torch.empty(n, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    en.record(); torch.cuda.synchronize()
    ms = st.elapsed_time(en) / 5
    print("pinned H2D %4d MB: %.2f ms = %.1f GB/s (is_pinned %s)" % (mb, ms, mb / 1024 / (ms / 1e3), h.is_pinned()))
    h2 = torch.empty(n).pin_memory()
    st.record()
    for _ in range(5):
        h2.copy_(d, non_blocking=True)
    en.record(); torch.cuda.synchronize()
    ms = st.elapsed_time(en) / 5
    print("pinned D2H %4d MB: %.2f ms = %.1f GB/s" % (mb, ms, mb / 1024 / (ms / 1e3)))
t = torch.randn(54 * 1024 * 1024 // 4)
for thr in (1, 4, 16):
    torch.set_num_threads(thr)
    o = torch.empty_like(t, dtype=torch.bfloat16).pin_memory()
    t0 = time.perf_counter()
    for _ in range(5):
        o.copy_(t)
    print("host fp32->bf16 54 MB, %d threads: %.2f ms" % (thr, (time.perf_counter() - t0) / 5 * 1e3))
