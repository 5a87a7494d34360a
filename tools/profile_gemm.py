"""Run the GEMM kernel at the step's main shapes/epilogues (for ncu and quick timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import ops
from x2vlm_b200._capi import ACT_GELU, ACT_GELU_BWD
dev = torch.device("cuda:0"); torch.manual_seed(0)
M, D, Dh = 17730, 768, 3072
x = torch.randn(M, D, device=dev).bfloat16(); w1 = torch.randn(Dh, D, device=dev).bfloat16(); w2 = torch.randn(D, Dh, device=dev).bfloat16()
b1 = torch.randn(Dh, device=dev); b2 = torch.randn(D, device=dev); gam = torch.randn(D, device=dev)
rs = torch.ones(90, device=dev); res = torch.randn(M, D, device=dev)
h = torch.empty(M, Dh, device=dev, dtype=torch.bfloat16); a = torch.empty_like(h); y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
o32 = torch.empty(M, D, device=dev); g = torch.randn(M, D, device=dev).bfloat16(); dh = torch.empty_like(h); dx = torch.empty_like(y)
gw = torch.zeros(D, Dh, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
cases = {
    "fc1 fwd bias+gelu+preact": lambda: ops.gemm(x, w1, M, Dh, D, bias=b1, act=ACT_GELU, preact_out=h, out_bf16=a),
    "fc2 fwd bias+ls+res f32": lambda: ops.gemm(a, w2, M, D, Dh, bias=b2, preact_out=y, gamma=gam, row_scale=rs, rows_per_scale=197, residual=res, out_f32=o32),
    "fc2 dgrad gelu_bwd": lambda: ops.gemm(g, w2, M, Dh, D, b_mn=True, act=ACT_GELU_BWD, aux=h, out_bf16=dh),
    "fc1 dgrad plain": lambda: ops.gemm(dh, w1, M, D, Dh, b_mn=True, out_bf16=dx),
    "fc2 wgrad splitk": lambda: ops.gemm(g, a, D, Dh, M, a_mn=True, b_mn=True, out_f32=gw, accumulate=True),
    "plain fwd bf16": lambda: ops.gemm(x, w1, M, Dh, D, out_bf16=a),
}
for name, fn in cases.items():
    for _ in range(2):
        fn()
    ts = []
    for cold in (False, True):
        t = []
        for _ in range(5):
            if cold:
                flush.zero_()
            st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.record(); fn(); en.record(); torch.cuda.synchronize(); t.append(st.elapsed_time(en))
        ts.append(sorted(t)[2])
    fl = 2.0 * M * D * Dh
    print("GEMM %-28s warm %.1f us (%.0f TF)  cold %.1f us (%.0f TF)" % (name, ts[0] * 1e3, fl / ts[0] / 1e9, ts[1] * 1e3, fl / ts[1] / 1e9))
