"""A/B timing of GEMM epilogue variants at the step's N = K = 768 projection shapes (developer tool, CUDA events, warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)


def bench(fn, reps=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record(); fn(); en.record(); torch.cuda.synchronize()
        tot += st.elapsed_time(en)
    return tot / reps * 1e3


for M, N, K in ((23040, 768, 768), (17730, 768, 768), (10240, 768, 768), (23040, 768, 3072)):
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    of = torch.empty(M, N, device=dev); ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    cases = {
        "bf16": dict(out_bf16=ob),
        "bias+bf16": dict(bias=bias, out_bf16=ob),
        "bias+drop+bf16": dict(bias=bias, dropout_p=0.1, dropout_seed=1, out_bf16=ob),
        "bias+f32": dict(bias=bias, out_f32=of),
        "bias+drop+f32": dict(bias=bias, dropout_p=0.1, dropout_seed=1, out_f32=of),
        "bias+res+f32": dict(bias=bias, residual=res, out_f32=of),
        "bias+drop+res+f32": dict(bias=bias, dropout_p=0.1, dropout_seed=1, residual=res, out_f32=of),
    }
    for name, kw in cases.items():
        us = bench(lambda: ops.gemm(a, w, M, N, K, **kw))
        print("GEMM %dx%dx%d %-20s %7.1f us  %6.0f TFLOP/s" % (M, N, K, name, us, 2.0 * M * N * K / us / 1e6))
