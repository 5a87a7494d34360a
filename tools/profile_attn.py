"""Run the attention kernels alone at the step's shapes (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
H, D = 12, 768
def step_kv_index():
    """kv_index of the step's fusion batch (pretrain.XVLM.forward_mixed): 576 sequences over 154 K/V sources."""
    g = torch.Generator().manual_seed(0)
    ar = torch.arange(64)
    img = [ar, torch.randint(0, 64, (64,), generator=g), ar, ar]
    reg = [64 + ar, 64 + torch.randint(0, 64, (64,), generator=g), 64 + ar, 64 + ar, 128 + torch.randint(0, 26, (64,), generator=g).sort().values]
    return torch.cat(img + reg).int().to(dev)


def run(name, B, Lq, Lk, n_kv, bias=False, mask=False, shared=False, p=0.0, reps=3, grouped=False):
    ld = ops.pad32(Lk)
    q = torch.randn(B * Lq, 3 * D, device=dev).bfloat16()
    kv = torch.randn(n_kv * Lk, 2 * D, device=dev).bfloat16()
    if Lq == Lk and not shared:
        qv, kk, vv = q[:, :D], q[:, D:2 * D], q[:, 2 * D:]
    else:
        qv, kk, vv = q[:, :D], kv[:, :D], kv[:, D:]
    b = torch.randn(H, Lq, ld, device=dev) if bias else None
    m = torch.zeros(B, ld, device=dev) if mask else None
    idx = (step_kv_index() if grouped else (torch.arange(B, device=dev) % n_kv).int()) if shared else None
    groups = ops.attn_group_table(idx, n_kv, Lq, Lk) if grouped else None
    o = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, Lq, device=dev)
    do = torch.randn(B * Lq, D, device=dev).bfloat16()
    dq = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16); dkv = torch.empty((n_kv if grouped else B) * Lk, 2 * D, device=dev, dtype=torch.bfloat16)
    ds = torch.empty(B, H, Lq, ld, device=dev, dtype=torch.bfloat16) if bias and not os.environ.get("X2K_ATTN_NO_DS") else None
    kw = dict(kv_index=idx, n_kv=n_kv, kv_groups=groups, bias=b, mask=m, dropout_p=p, dropout_seed=1, dropout_offset=0)
    for _ in range(reps):
        ops.attn_fwd(qv, kk, vv, B, H, Lq, Lk, 0.125, o, lse, **kw)
        ops.attn_bwd(qv, kk, vv, B, H, Lq, Lk, 0.125, o, lse, do, dq, dkv[:, :D], dkv[:, D:], ds_out=ds, **kw)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record(); ops.attn_fwd(qv, kk, vv, B, H, Lq, Lk, 0.125, o, lse, **kw); en.record(); torch.cuda.synchronize(); tf = st.elapsed_time(en)
    st.record(); ops.attn_bwd(qv, kk, vv, B, H, Lq, Lk, 0.125, o, lse, do, dq, dkv[:, :D], dkv[:, D:], ds_out=ds, **kw); en.record(); torch.cuda.synchronize()
    print("ATTN %-12s B=%d Lq=%d Lk=%d fwd %.1f us bwd %.1f us" % (name, B, Lq, Lk, tf * 1e3, st.elapsed_time(en) * 1e3))
only = os.environ.get("X2K_ATTN_CASE")
for case in (("beit", 90, 197, 197, 90, dict(bias=True)), ("beit-nobias", 90, 197, 197, 90, dict()), ("beit577", 16, 577, 577, 16, dict(bias=True)),
             ("beit2305", 2, 2305, 2305, 2, dict(bias=True)), ("cross577", 64, 40, 577, 16, dict(mask=True, shared=True, p=0.1)),
             ("text", 256, 40, 40, 256, dict(mask=True, p=0.1)),
             ("fus-self", 576, 40, 40, 576, dict(mask=True, p=0.1)),
             ("cross", 576, 40, 197, 154, dict(mask=True, shared=True, p=0.1, grouped=True))):
    if only is None or only == case[0]:
        run(*case[:5], **case[5])
