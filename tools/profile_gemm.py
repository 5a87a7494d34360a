"""Run the GEMM kernel at the step's main shapes/epilogues (for ncu and quick timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import ops
from x2vlm_b200._capi import ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX
dev = torch.device("cuda:0"); torch.manual_seed(0)
M, D, Dh = 17730, 768, 3072
x = torch.randn(M, D, device=dev).bfloat16(); w1 = torch.randn(Dh, D, device=dev).bfloat16(); w2 = torch.randn(D, Dh, device=dev).bfloat16()
b1 = torch.randn(Dh, device=dev); b2 = torch.randn(D, device=dev); gam = torch.randn(D, device=dev)
rs = torch.ones(90, device=dev); res = torch.randn(M, D, device=dev)
h = torch.empty(M, Dh, device=dev, dtype=torch.bfloat16); a = torch.empty_like(h); y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
o32 = torch.empty(M, D, device=dev); g = torch.randn(M, D, device=dev).bfloat16(); dh = torch.empty_like(h); dx = torch.empty_like(y)
gw = torch.zeros(D, Dh, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
M2 = 23040
x2 = torch.randn(M2, D, device=dev).bfloat16(); wo = torch.randn(D, D, device=dev).bfloat16(); res2 = torch.randn(M2, D, device=dev)
o2 = torch.empty(M2, D, device=dev); o2b = torch.empty(M2, D, device=dev, dtype=torch.bfloat16)
small = {  # attention output projections: N = K = 768 (flops differ from the MLP shapes; reported separately)
    "o-proj bias+f32": lambda: ops.gemm(x2, wo, M2, D, D, bias=b2, out_f32=o2),
    "o-proj bias+bf16": lambda: ops.gemm(x2, wo, M2, D, D, bias=b2, out_bf16=o2b),
    "o-proj bias+drop+res+f32": lambda: ops.gemm(x2, wo, M2, D, D, bias=b2, dropout_p=0.1, dropout_seed=1, dropout_offset=0, residual=res2, out_f32=o2),
    "o-proj bias+res+f32": lambda: ops.gemm(x2, wo, M2, D, D, bias=b2, residual=res2, out_f32=o2),
}
cases = {
    "fc1 fwd bias+gelu+gelu'": lambda: ops.gemm(x, w1, M, Dh, D, bias=b1, act=ACT_GELU_SAVE_GRAD, preact_out=h, out_bf16=a),
    "fc1 fwd bias only": lambda: ops.gemm(x, w1, M, Dh, D, bias=b1, out_bf16=a),
    "fc2 fwd bias+ls+res f32": lambda: ops.gemm(a, w2, M, D, Dh, bias=b2, preact_out=y, gamma=gam, row_scale=rs, rows_per_scale=197, residual=res, out_f32=o32),
    "fc2 dgrad mul_aux": lambda: ops.gemm(g, w2, M, Dh, D, b_mn=True, act=ACT_MUL_AUX, aux=h, out_bf16=dh),
    "dense+drop+res f32": lambda: ops.gemm(a, w2, M, D, Dh, bias=b2, dropout_p=0.1, dropout_seed=1, dropout_offset=0, residual=res, out_f32=o32),
    "fc1 dgrad plain": lambda: ops.gemm(dh, w1, M, D, Dh, b_mn=True, out_bf16=dx),
    "fc2 wgrad splitk": lambda: ops.gemm(g, a, D, Dh, M, a_mn=True, b_mn=True, out_f32=gw, accumulate=True),
    "plain fwd bf16": lambda: ops.gemm(x, w1, M, Dh, D, out_bf16=a),
}
only = os.environ.get("X2K_CASE")
allc = [(n, f, 2.0 * M * D * Dh) for n, f in cases.items()] + [(n, f, 2.0 * M2 * D * D) for n, f in small.items()]
for name, fn, fl in allc:
    if only and only != name:
        continue
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
    print("GEMM %-28s warm %.1f us (%.0f TF)  cold %.1f us (%.0f TF)" % (name, ts[0] * 1e3, fl / ts[0] / 1e9, ts[1] * 1e3, fl / ts[1] / 1e9))
