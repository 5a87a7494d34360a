"""GPU parity tests of the C-ABI kernels (called through x2vlm_b200.ops -> libx2k.so) against fp32 torch
restatements of the cited reference lines and the numpy Philox oracle.  Tolerances: operands are bf16, all
accumulation is fp32, so results are compared with the fp32 reference evaluated on the SAME bf16-rounded
inputs; bound = a few bf16 ulps of the output magnitude (stated per test)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _bf(x):
    return x.bfloat16()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 328, 200), (12608, 768, 768), (400, 2304, 768)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_majors(dev, M, N, K, a_mn, b_mn):
    from x2vlm_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    A = _bf(torch.randn(M, K, device=dev, generator=g)); B = _bf(torch.randn(N, K, device=dev, generator=g))
    ref = A.float() @ B.float().t()
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    for tn in (128, 256):
        out = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn, out_f32=out, tile_n=tn)
        # fp32 accumulation of exact bf16 products: only summation-order error remains
        assert (out - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() * math.sqrt(K / 64)


def test_gemm_split_k_wgrad(dev):
    """wgrad-shaped GEMM (small output, K = tokens): split-K with atomic fp32 accumulation."""
    from x2vlm_b200 import ops
    g = torch.Generator(device=dev).manual_seed(7)
    rows, n_out, n_in = 17730, 768, 768
    dy = _bf(torch.randn(rows, n_out, device=dev, generator=g)); x = _bf(torch.randn(rows, n_in, device=dev, generator=g))
    ref = dy.float().t() @ x.float()
    base = torch.randn(n_out, n_in, device=dev, generator=g)
    for split in (0, 1, 5):
        out = base.clone()
        ops.gemm(dy, x, n_out, n_in, rows, a_mn=True, b_mn=True, out_f32=out, accumulate=True, split_k=split)
        assert (out - (base + ref)).abs().max().item() < 2e-3 * ref.abs().max().item(), split
        out2 = torch.full((n_out, n_in), float("nan"), device=dev)
        ops.gemm(dy, x, n_out, n_in, rows, a_mn=True, b_mn=True, out_f32=out2, split_k=split)
        assert (out2 - ref).abs().max().item() < 2e-3 * ref.abs().max().item(), split


def test_gemm_epilogues(dev):
    from x2vlm_b200 import ops
    from x2vlm_b200._capi import ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX
    from oracle import philox
    M, N, K = 394, 768, 768
    g = torch.Generator(device=dev).manual_seed(0)
    A = _bf(torch.randn(M, K, device=dev, generator=g)); W = _bf(torch.randn(N, K, device=dev, generator=g) * 0.05)
    bias = torch.randn(N, device=dev, generator=g); gamma = torch.randn(N, device=dev, generator=g)
    res = torch.randn(M, N, device=dev, generator=g); rs = torch.rand(2, device=dev, generator=g) + 0.5
    acc = A.float() @ W.float().t()
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(o)
    ops.gemm(A, W, M, N, K, bias=bias, act=ACT_GELU, preact_out=pre, out_bf16=o)
    assert (pre.float() - (acc + bias)).abs().max() < 0.04  # bf16 rounding of |x| <= 8
    assert (o.float() - torch.nn.functional.gelu(acc + bias)).abs().max() < 0.04
    o32 = torch.empty(M, N, device=dev)
    ops.gemm(A, W, M, N, K, bias=bias, gamma=gamma, row_scale=rs, rows_per_scale=197, residual=res, out_f32=o32)
    ref = res + (acc + bias) * gamma * rs.repeat_interleave(197)[:, None]
    assert (o32 - ref).abs().max() < 1e-4 * ref.abs().max()
    h = _bf(torch.randn(M, N, device=dev, generator=g))
    ops.gemm(A, W, M, N, K, act=ACT_GELU_BWD, aux=h, out_bf16=o)
    hf = h.float().requires_grad_(True); torch.nn.functional.gelu(hf).sum().backward()
    assert (o.float() - acc * hf.grad).abs().max() < 0.02 * (acc * hf.grad).abs().max()
    # the hot-path pair: forward stores GELU'(pre) next to GELU(pre), backward multiplies by the stored derivative
    dg = torch.empty_like(o)
    ops.gemm(A, W, M, N, K, bias=bias, act=ACT_GELU_SAVE_GRAD, preact_out=dg, out_bf16=o)
    xf = (acc + bias).clone().requires_grad_(True); torch.nn.functional.gelu(xf).sum().backward()
    assert (o.float() - torch.nn.functional.gelu(acc + bias)).abs().max() < 0.04
    assert (dg.float() - xf.grad).abs().max() < 6e-3          # |GELU'| <= 1.13: half a bf16 ulp is 3.9e-3
    o2 = torch.empty_like(o)
    ops.gemm(A, W, M, N, K, act=ACT_MUL_AUX, aux=h, out_bf16=o2)
    assert (o2.float() - acc * h.float()).abs().max() < 0.01 * (acc * h.float()).abs().max()
    # fp32 accuracy of the GELU formulas themselves (large-|x| tails included), through an fp32 output
    xs = torch.linspace(-9, 9, M * N, device=dev).view(M, N)
    eye_a = torch.zeros(M, K, device=dev); eye_w = torch.zeros(N, K, device=dev)   # acc == 0, bias carries x per column only
    xs_col = torch.linspace(-9, 9, N, device=dev)
    ops.gemm(_bf(eye_a), _bf(eye_w), M, N, K, bias=xs_col, act=ACT_GELU, out_f32=o32)
    assert (o32[0] - torch.nn.functional.gelu(xs_col.double()).float()).abs().max() < 2e-6
    o32b = o32.clone()
    ops.gemm(A, W, M, N, K, accumulate=True, out_f32=o32b)
    assert (o32b - (o32 + acc)).abs().max() < 1e-4 * o32.abs().max()
    # dropout: the exact Philox mask of the oracle
    ops.gemm(A, W, M, N, K, dropout_p=0.1, dropout_seed=1234, dropout_offset=77, out_f32=o32)
    keep = torch.from_numpy(philox.keep_scale(1234, 77, M * N, 0.1)).view(M, N).to(dev)
    assert (o32 - acc * keep).abs().max() < 1e-4 * acc.abs().max()


@pytest.mark.parametrize("M,D,eps", [(12608, 768, 1e-6), (2560, 768, 1e-12), (333, 1024, 1e-6), (77, 128, 1e-5)])
def test_layernorm(dev, M, D, eps):
    from x2vlm_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M)
    x = torch.randn(M, D, device=dev, generator=g) * 2 + 0.5
    w = torch.randn(D, device=dev, generator=g); b = torch.randn(D, device=dev, generator=g)
    yb = torch.empty(M, D, device=dev, dtype=torch.bfloat16); yf = torch.empty(M, D, device=dev)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    ops.layernorm_fwd(x, w, b, eps, y_bf16=yb, y_f32=yf, mean=mean, rstd=rstd)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), wr, br, eps)
    assert (yf - ref).abs().max() < 1e-4 and (yb.float() - ref).abs().max() < 0.05
    dy = torch.randn(M, D, device=dev, generator=g); dyb = _bf(torch.randn(M, D, device=dev, generator=g))
    res = torch.randn(M, D, device=dev, generator=g)
    ref.backward(dy + dyb.float())
    dx = torch.empty(M, D, device=dev); dw = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    ops.layernorm_bwd((dy, dyb), x, w, mean, rstd, dx, dw, db, dx_residual=res)
    assert (dx - (xr.grad + res)).abs().max() < 1e-3
    assert (dw - wr.grad).abs().max() < 1e-3 * wr.grad.abs().max() and (db - br.grad).abs().max() < 1e-3 * br.grad.abs().max()
    # fused variant: same dx/dw/db plus g = bf16(dx * Philox keep-scale) and dbias += colsum(dx * keep-scale)
    from oracle import philox
    dx2 = torch.empty(M, D, device=dev); dw2 = torch.zeros(D, device=dev); db2 = torch.zeros(D, device=dev)
    gb = torch.empty(M, D, device=dev, dtype=torch.bfloat16); dbias = torch.zeros(D, device=dev)
    ops.layernorm_bwd((dy, dyb), x, w, mean, rstd, dx2, dw2, db2, dx_residual=res, g_bf16=gb, dbias=dbias, dropout_p=0.1,
                      dropout_seed=11, dropout_offset=5)
    assert (dx2 - dx).abs().max() <= 1e-5 * dx.abs().max()       # same math; FMA contraction may differ between the two instantiations
    assert (dw2 - dw).abs().max() <= 1e-4 * dw.abs().max() and (db2 - db).abs().max() <= 1e-4 * db.abs().max()
    keep = torch.from_numpy(philox.keep_scale(11, 5, M * D, 0.1)).view(M, D).to(dev)
    assert (gb.float() - dx * keep).abs().max() < 0.01 * (dx * keep).abs().max()
    assert (dbias - (dx * keep).sum(0)).abs().max() < 1e-3 * (dx * keep).sum(0).abs().max() + 1e-3
    ops.layernorm_bwd((dy, dyb), x, w, mean, rstd, dx2, dw2, db2, g_bf16=gb, dbias=dbias)      # p = 0: plain cast
    assert (gb.float() - dx2).abs().max() < 0.01 * dx2.abs().max()


def test_scale_cast_colsum_and_friends(dev):
    from x2vlm_b200 import ops
    from oracle import philox
    M, N = 1970, 768
    g = torch.Generator(device=dev).manual_seed(3)
    dx = torch.randn(M, N, device=dev, generator=g); gamma = torch.randn(N, device=dev, generator=g)
    rs = torch.rand(10, device=dev, generator=g) + 0.5; y = _bf(torch.randn(M, N, device=dev, generator=g))
    gb = torch.empty(M, N, device=dev, dtype=torch.bfloat16); dbias = torch.zeros(N, device=dev); dgamma = torch.zeros(N, device=dev)
    ops.scale_cast_colsum(dx, M, N, g_bf16=gb, gamma=gamma, row_scale=rs, rows_per_scale=197, dropout_p=0.1, dropout_seed=5,
                          dropout_offset=9, y_bf16=y, dbias=dbias, dgamma=dgamma)
    keep = torch.from_numpy(philox.keep_scale(5, 9, M * N, 0.1)).view(M, N).to(dev)
    rsx = rs.repeat_interleave(197)[:, None]
    gref = dx * rsx * gamma * keep
    assert (gb.float() - gref).abs().max() <= 2 ** -8 * gref.abs().max()  # one bf16 rounding
    assert (dbias - gref.sum(0)).abs().max() < 1e-3 * gref.sum(0).abs().max()
    dgref = (dx * rsx * y.float()).sum(0)
    assert (dgamma - dgref).abs().max() < 1e-3 * dgref.abs().max()
    xb = _bf(torch.randn(5120, 2304, device=dev, generator=g)); dc = torch.zeros(2304, device=dev)
    ops.colsum_bf16(xb, 5120, 2304, dc)
    assert (dc - xb.float().sum(0)).abs().max() < 1e-3 * xb.float().sum(0).abs().max()
    src = _bf(torch.randn(11, 197 * 64, device=dev, generator=g)); idx = torch.tensor([2, 0, 2, 1, 1, 2, 0, 0, 3, 2, 1], device=dev, dtype=torch.int32)
    out = torch.empty(5, 197 * 64, device=dev, dtype=torch.bfloat16)
    ops.segment_sum_bf16(src, idx, 5, out)
    ref = torch.zeros(5, 197 * 64, device=dev).index_add_(0, idx.long(), src.float())
    assert (out.float() - ref).abs().max() <= 2 ** -7 * ref.abs().max() and out[4].abs().max() == 0


def _attn_ref(q, k, v, scale, bias=None, mask=None, keep=None):
    """fp32 restatement of models/beit2.py:135-159 / models/xbert.py:364-410 on [B,H,L,64] tensors."""
    s = (q @ k.transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias
    if mask is not None:
        s = s + mask
    p = s.softmax(-1)
    pd = p if keep is None else p * keep
    return pd @ v, p


def _heads(x, B, L, H):
    return x.view(B, L, H, 64).permute(0, 2, 1, 3)


@pytest.mark.parametrize("case", ["beit", "text", "text3d", "cross_shared", "text_dropout", "n256",
                                  # whole-range kernels (attn.cu) beyond the BEiT shape: one tile with key mask + dropout,
                                  # 2 x 2 tiles with a per-query mask + dropout, two query tiles over one shared key block
                                  "self100_dropout", "causal200_dropout", "q200_k72_shared",
                                  # key-blocked online-softmax kernels (attn_long.cu): 384 px / 768 px fine-tuning shapes
                                  "beit577", "beit2305", "cross577_shared_dropout", "causal300", "cross2305"])
def test_attention_fwd_bwd(dev, case):
    from x2vlm_b200 import ops
    from oracle import philox
    g = torch.Generator(device=dev).manual_seed(11)
    H = 12
    kv_index = None
    p_drop = 0.0
    if case == "beit":
        B, Lq, Lk, n_kv = 3, 197, 197, 3
    elif case == "n256":
        B, Lq, Lk, n_kv, H = 2, 256, 256, 2, 2
    elif case == "cross_shared":
        B, Lq, Lk, n_kv = 7, 40, 197, 3
        kv_index = torch.tensor([0, 2, 1, 1, 0, 2, 2], device=dev, dtype=torch.int32)
    elif case == "self100_dropout":
        B, Lq, Lk, n_kv, H = 3, 100, 100, 3, 4
        p_drop = 0.1
    elif case == "causal200_dropout":
        B, Lq, Lk, n_kv, H = 2, 200, 200, 2, 4
        p_drop = 0.1
    elif case == "q200_k72_shared":
        B, Lq, Lk, n_kv, H = 5, 200, 72, 2, 4
        kv_index = torch.tensor([1, 0, 0, 1, 1], device=dev, dtype=torch.int32)
    elif case == "beit577":      # 384 px: 24 x 24 patches + cls, dense rel-pos bias streamed per key block
        B, Lq, Lk, n_kv, H = 2, 577, 577, 2, 2
    elif case == "beit2305":     # 768 px: 48 x 48 patches + cls (configs/finetune/vqa2_base.yaml)
        B, Lq, Lk, n_kv, H = 1, 2305, 2305, 1, 2
    elif case == "cross577_shared_dropout":  # captions over 384 px image tokens, shared K/V, key mask, prob. dropout
        B, Lq, Lk, n_kv, H = 7, 40, 577, 3, 4
        kv_index = torch.tensor([0, 2, 1, 1, 0, 2, 2], device=dev, dtype=torch.int32)
        p_drop = 0.1
    elif case == "cross2305":
        B, Lq, Lk, n_kv, H = 2, 40, 2305, 2, 2
    elif case == "causal300":    # more than two query tiles with a per-query (3-D) mask
        B, Lq, Lk, n_kv, H = 2, 300, 300, 2, 2
    else:
        B, Lq, Lk, n_kv = 5, 40, 40, 5
    if case == "text_dropout":
        p_drop = 0.1
    D = H * 64
    ld = ops.pad32(Lk)      # row stride of bias / mask / dS buffers
    lkp = ops.pad16(Lk)     # the kernels' padded key count (dropout element index)
    scale = 0.125
    qkv = _bf(torch.randn(B * Lq, 3 * D, device=dev, generator=g))
    if Lq == Lk and kv_index is None:
        qv, kv_, vv = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        kvbuf = _bf(torch.randn(n_kv * Lk, 2 * D, device=dev, generator=g))
        qv, kv_, vv = qkv[:, :D], kvbuf[:, :D], kvbuf[:, D:]
    bias = mask = None
    per_query = False
    if case in ("beit", "n256", "beit577", "beit2305"):
        bias = torch.zeros(H, Lq, ld, device=dev); bias[:, :, :Lk] = torch.randn(H, Lq, Lk, device=dev, generator=g)
    if case in ("text", "text_dropout", "cross_shared", "cross577_shared_dropout", "cross2305", "self100_dropout", "q200_k72_shared"):
        m01 = (torch.rand(B, Lk, device=dev, generator=g) > 0.3).float(); m01[:, 0] = 1
        mask = torch.zeros(B, ld, device=dev); mask[:, :Lk] = (1 - m01) * -10000.0
    if case in ("text3d", "causal300", "causal200_dropout"):
        m01 = torch.tril(torch.ones(Lq, Lk, device=dev)).expand(B, -1, -1)
        mask = torch.zeros(B, Lq, ld, device=dev); mask[:, :, :Lk] = (1 - m01) * -10000.0
        per_query = True
    o = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, Lq, device=dev)
    kw = dict(kv_index=kv_index, n_kv=n_kv, bias=bias, mask=mask, mask_per_query=per_query, dropout_p=p_drop, dropout_seed=42,
              dropout_offset=1000)
    ops.attn_fwd(qv, kv_, vv, B, H, Lq, Lk, scale, o, lse, **kw)
    # ---- fp32 reference on the same bf16 inputs ----
    sel = kv_index.long() if kv_index is not None else torch.arange(B, device=dev)
    qf = _heads(qv.float(), B, Lq, H).requires_grad_(True)
    kf = _heads(kv_.float().reshape(n_kv, Lk, D)[sel].reshape(B * Lk, D), B, Lk, H).detach().requires_grad_(True)
    vf = _heads(vv.float().reshape(n_kv, Lk, D)[sel].reshape(B * Lk, D), B, Lk, H).detach().requires_grad_(True)
    bias_r = bias[:, :, :Lk].unsqueeze(0).clone().requires_grad_(True) if bias is not None else None
    mask_r = None
    if mask is not None:
        mask_r = mask[:, None, :, :Lk] if per_query else mask[:, None, None, :Lk]
    keep = None
    if p_drop > 0:
        keep = torch.from_numpy(philox.keep_scale(42, 1000, B * H * Lq * lkp, p_drop)).view(B, H, Lq, lkp)[..., :Lk].to(dev)
    oref, pref = _attn_ref(qf, kf, vf, scale, bias_r, mask_r, keep)
    o_cmp = _heads(o.float(), B, Lq, H)
    tol = 2e-2  # P is rounded to bf16 before P·V: relative error ~2^-8 per term, outputs are O(1)
    assert (o_cmp - oref).abs().max().item() < tol * max(1.0, oref.abs().max().item())
    sref = (qf @ kf.transpose(-1, -2)) * scale + (bias_r if bias_r is not None else 0) + (mask_r if mask_r is not None else 0)
    lse_ref = torch.logsumexp(sref, -1) / math.log(2.0)
    assert (lse - lse_ref).abs().max().item() < 2e-3
    # ---- backward ----
    do = _bf(torch.randn(B * Lq, D, device=dev, generator=g))
    dq = torch.full((B * Lq, D), float("nan"), device=dev, dtype=torch.bfloat16)
    dkv = torch.full((B * Lk, 2 * D), float("nan"), device=dev, dtype=torch.bfloat16)
    ds = torch.zeros(B, H, Lq, ld, device=dev, dtype=torch.bfloat16) if bias is not None else None
    ops.attn_bwd(qv, kv_, vv, B, H, Lq, Lk, scale, o, lse, do, dq, dkv[:, :D], dkv[:, D:], ds_out=ds, **kw)
    oref.backward(_heads(do.float(), B, Lq, H))
    for name, got, want in (("dq", _heads(dq.float(), B, Lq, H), qf.grad), ("dk", _heads(dkv[:, :D].float(), B, Lk, H), kf.grad),
                            ("dv", _heads(dkv[:, D:].float(), B, Lk, H), vf.grad)):
        assert not torch.isnan(got).any(), name
        err = (got - want).abs().max().item()
        assert err < 3e-2 * max(1.0, want.abs().max().item()), (name, err, want.abs().max().item())
    if bias is not None:
        dbias = ds.float().sum(0)[:, :, :Lk]
        want = bias_r.grad[0]
        assert (dbias - want).abs().max().item() < 3e-2 * max(1.0, want.abs().max().item())


def _group_table_ref(kv, n_kv, G):
    """Python restatement of x2k_attn_group_build: sequences grouped by source in batch order, G per item."""
    by_src = [[] for _ in range(n_kv)]
    for b, s_ in enumerate(kv):
        by_src[s_].append(b)
    first, items = [0], []
    for s_ in range(n_kv):
        seqs = by_src[s_]
        for i in range(0, len(seqs), G):
            chunk = seqs[i:i + G]
            items.append([s_, len(chunk)] + chunk + [-1] * (8 - len(chunk)) + [0, 0])
        first.append(len(items))
    return first, items


@pytest.mark.parametrize("case", ["cross_grouped", "cross_grouped_l30", "cross_identity", "self_l30", "self_l16_3d", "self_l64"])
def test_attention_packed(dev, case):
    """Packed short-sequence kernels (attn_pack.cu): several sequences per MMA tile; grouped cross-attention returns
    dK/dV per K/V source (summed over the sequences sharing it)."""
    from x2vlm_b200 import ops
    from x2vlm_b200 import _capi as C
    from oracle import philox
    g = torch.Generator(device=dev).manual_seed(23)
    H, scale = 4, 0.125
    D = H * 64
    kv_list = None
    p_drop, use_mask, per_query = 0.0, False, False
    if case == "cross_grouped":      # source 1 unused, source 2 needs two work items (5 sequences, G = 3)
        B, Lq, Lk, n_kv = 11, 40, 197, 4
        kv_list = [0, 2, 3, 2, 0, 2, 2, 3, 2, 0, 3]
        p_drop, use_mask = 0.1, True
    elif case == "cross_grouped_l30":
        B, Lq, Lk, n_kv = 9, 30, 50, 3
        kv_list = [2, 2, 0, 1, 2, 2, 2, 0, 1]
    elif case == "cross_identity":
        B, Lq, Lk, n_kv = 5, 40, 197, 5
        kv_list = list(range(5))
        use_mask = True
    elif case == "self_l30":
        B, Lq, Lk, n_kv = 9, 30, 30, 9
        p_drop, use_mask = 0.1, True
    elif case == "self_l16_3d":
        B, Lq, Lk, n_kv = 19, 16, 16, 19
        use_mask = per_query = True
    else:
        B, Lq, Lk, n_kv = 5, 64, 64, 5
        use_mask = True
    cross = kv_list is not None
    ld, lkp = ops.pad32(Lk), ops.pad16(Lk)
    qkv = _bf(torch.randn(B * Lq, 3 * D, device=dev, generator=g))
    if cross:
        kvbuf = _bf(torch.randn(n_kv * Lk, 2 * D, device=dev, generator=g))
        qv, kv_, vv = qkv[:, :D], kvbuf[:, :D], kvbuf[:, D:]
        kv_index = torch.tensor(kv_list, device=dev, dtype=torch.int32)
        table = ops.attn_group_table(kv_index, n_kv, Lq, Lk)
        assert table is not None
        G = C.lib().x2k_attn_group_slots(Lq)
        first, items = _group_table_ref(kv_list, n_kv, G)
        t = table.cpu().tolist()
        assert t[:4] == [len(items), G, n_kv, B]
        assert t[4:4 + n_kv + 1] == first
        off = 4 + ((n_kv + 1 + 3) // 4) * 4
        for i, it in enumerate(items):
            assert t[off + 12 * i: off + 12 * i + 10] == it[:10], (i, t[off + 12 * i: off + 12 * i + 12], it)
    else:
        qv, kv_, vv = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        kv_index = table = None
    mask = None
    if use_mask and not per_query:
        m01 = (torch.rand(B, Lk, device=dev, generator=g) > 0.3).float(); m01[:, 0] = 1
        mask = torch.zeros(B, ld, device=dev); mask[:, :Lk] = (1 - m01) * -10000.0
    if per_query:
        m01 = torch.tril(torch.ones(Lq, Lk, device=dev)).expand(B, -1, -1)
        mask = torch.zeros(B, Lq, ld, device=dev); mask[:, :, :Lk] = (1 - m01) * -10000.0
    o = torch.full((B * Lq, D), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.full((B, H, Lq), float("nan"), device=dev)
    kw = dict(kv_index=kv_index, n_kv=n_kv, kv_groups=table, mask=mask, mask_per_query=per_query, dropout_p=p_drop,
              dropout_seed=77, dropout_offset=4096)
    ops.attn_fwd(qv, kv_, vv, B, H, Lq, Lk, scale, o, lse, **kw)
    sel = kv_index.long() if cross else torch.arange(B, device=dev)
    qf = _heads(qv.float(), B, Lq, H).requires_grad_(True)
    k_src = _heads(kv_.float(), n_kv, Lk, H).detach().requires_grad_(True)   # [n_kv, H, Lk, 64]
    v_src = _heads(vv.float(), n_kv, Lk, H).detach().requires_grad_(True)
    kf, vf = k_src[sel], v_src[sel]
    mask_r = None
    if mask is not None:
        mask_r = mask[:, None, :, :Lk] if per_query else mask[:, None, None, :Lk]
    keep = None
    if p_drop > 0:
        keep = torch.from_numpy(philox.keep_scale(77, 4096, B * H * Lq * lkp, p_drop)).view(B, H, Lq, lkp)[..., :Lk].to(dev)
    oref, _ = _attn_ref(qf, kf, vf, scale, None, mask_r, keep)
    o_cmp = _heads(o.float(), B, Lq, H)
    assert not torch.isnan(o_cmp).any() and not torch.isnan(lse).any()
    assert (o_cmp - oref).abs().max().item() < 2e-2 * max(1.0, oref.abs().max().item())
    sref = (qf @ kf.transpose(-1, -2)) * scale + (mask_r if mask_r is not None else 0)
    assert (lse - torch.logsumexp(sref, -1) / math.log(2.0)).abs().max().item() < 2e-3
    do = _bf(torch.randn(B * Lq, D, device=dev, generator=g))
    dq = torch.full((B * Lq, D), float("nan"), device=dev, dtype=torch.bfloat16)
    dkv = torch.full((n_kv * Lk, 2 * D), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.attn_bwd(qv, kv_, vv, B, H, Lq, Lk, scale, o, lse, do, dq, dkv[:, :D], dkv[:, D:], **kw)
    oref.backward(_heads(do.float(), B, Lq, H))
    for name, got, want in (("dq", _heads(dq.float(), B, Lq, H), qf.grad), ("dk", _heads(dkv[:, :D].float(), n_kv, Lk, H), k_src.grad),
                            ("dv", _heads(dkv[:, D:].float(), n_kv, Lk, H), v_src.grad)):
        assert not torch.isnan(got).any(), name
        err = (got - want).abs().max().item()
        assert err < 3e-2 * max(1.0, want.abs().max().item()), (name, err, want.abs().max().item())


def test_relpos_gather_scatter(dev):
    from x2vlm_b200 import ops
    from oracle import restate
    N, H, R = 197, 12, 732
    g = torch.Generator(device=dev).manual_seed(2)
    table = torch.randn(R, H, device=dev, generator=g)
    index = restate.relative_position_index((14, 14)).to(dev)
    out = torch.empty(H, N, 224, device=dev)
    ops.relpos_bias_gather(table, index, N, H, out)
    ref = table[index.view(-1)].view(N, N, H).permute(2, 0, 1)
    assert torch.equal(out[:, :, :N], ref) and out[:, :, N:].abs().max() == 0
    for nb in (4, 5):  # the vectorised kernel walks the batch four at a time + a tail
        ds = _bf(torch.randn(nb, H, N, 224, device=dev, generator=g))
        ds[..., 208:] = float("nan")  # columns past pad16(N) are never written by the attention kernel, never read here
        dt = torch.zeros(R, H, device=dev)
        ops.relpos_bias_scatter(ds, nb, H, N, index, dt)
        want = torch.zeros(R, H, device=dev).index_add_(0, index.view(-1),
                                                        ds.float().sum(0)[:, :, :N].permute(1, 2, 0).reshape(N * N, H))
        assert (dt - want).abs().max() < 1e-3 * want.abs().max()
    # unaligned view (row stride not a multiple of 8): scalar fallback
    ds = _bf(torch.randn(3, H, N, 203, device=dev, generator=g))
    dt = torch.zeros(R, H, device=dev)
    ops.relpos_bias_scatter(ds, 3, H, N, index, dt)
    want = torch.zeros(R, H, device=dev).index_add_(0, index.view(-1), ds.float().sum(0)[:, :, :N].permute(1, 2, 0).reshape(N * N, H))
    assert (dt - want).abs().max() < 1e-3 * want.abs().max()


def test_flat_adamw_matches_torch(dev):
    from x2vlm_b200 import ops
    n = 100000
    g = torch.Generator(device=dev).manual_seed(4)
    p = torch.randn(n, device=dev, generator=g); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    pb = torch.empty(n, device=dev, dtype=torch.bfloat16)
    cuts = [0, 30000, 70000, n]; lrs = [1e-3, 2e-3, 5e-4]; wds = [0.01, 0.0, 0.05]
    prefs = [p[a:b].clone().requires_grad_(True) for a, b in zip(cuts[:-1], cuts[1:])]
    opt = torch.optim.AdamW([{"params": [pp], "lr": lr, "weight_decay": wd} for pp, lr, wd in zip(prefs, lrs, wds)],
                            betas=(0.9, 0.98), eps=1e-8)
    seg_end = torch.tensor(cuts[1:], device=dev, dtype=torch.int64)
    seg_lr = torch.tensor(lrs, device=dev); seg_wd = torch.tensor(wds, device=dev); gs = torch.tensor([0.5], device=dev)
    step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
    for _ in range(3):
        gr = torch.randn(n, device=dev, generator=g)
        for pp, a, b in zip(prefs, cuts[:-1], cuts[1:]):
            pp.grad = gr[a:b].clone() * 0.5
        opt.step()
        step_dev += 1
        ops.adamw_flat(p, gr, m, v, pb, n, seg_end, seg_lr, seg_wd, 0.9, 0.98, 1e-8, step_dev=step_dev, grad_scale=gs)
    assert (p - torch.cat([pp.detach() for pp in prefs])).abs().max() < 1e-5
    assert torch.equal(pb, p.bfloat16())
    out = torch.zeros(1, device=dev); ops.sumsq(p, out)
    assert abs(out.item() - (p.double() ** 2).sum().item()) < 1e-4 * out.item()


def test_attention_workspace_query_and_loud_failure(dev):
    """x2k_attn_bwd_workspace_bytes: 0 for the whole-range kernels (Lq, Lk <= 256), B*Lq*H*64 fp32 for the key-blocked ones;
    a long backward without the workspace fails loudly (X2K_ERR_ARG) instead of truncating or falling back."""
    import ctypes
    from x2vlm_b200 import _capi as C
    a = C.X2kAttnArgs()
    a.B, a.H, a.Lq, a.Lk = 3, 12, 197, 197
    assert C.lib().x2k_attn_bwd_workspace_bytes(ctypes.byref(a)) == 0
    a.Lq, a.Lk = 40, 577
    assert C.lib().x2k_attn_bwd_workspace_bytes(ctypes.byref(a)) == 3 * 40 * 12 * 64 * 4
    a.Lq, a.Lk = 577, 577
    assert C.lib().x2k_attn_bwd_workspace_bytes(ctypes.byref(a)) == 3 * 577 * 12 * 64 * 4
    B, H, L = 1, 2, 300
    t = torch.randn(B * L, 3 * H * 64, device=dev).bfloat16()
    o = torch.empty(B * L, H * 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, L, device=dev); delta = torch.empty(B, H, L, device=dev)
    dqkv = torch.empty_like(t)
    a = C.X2kAttnArgs()
    a.q, a.k, a.v = t.data_ptr(), t.data_ptr() + 2 * H * 64, t.data_ptr() + 4 * H * 64
    a.ld_q = a.ld_k = a.ld_v = a.ld_dq = a.ld_dk = a.ld_dv = 3 * H * 64
    a.B, a.H, a.Lq, a.Lk, a.scale = B, H, L, L, 0.125
    a.o, a.ld_o, a.lse, a.d_o, a.ld_do = o.data_ptr(), H * 64, lse.data_ptr(), o.data_ptr(), H * 64
    a.dq, a.dk, a.dv = dqkv.data_ptr(), dqkv.data_ptr() + 2 * H * 64, dqkv.data_ptr() + 4 * H * 64
    a.delta_ws = delta.data_ptr()
    rc = C.lib().x2k_attn_bwd(ctypes.byref(a), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == -1 and b"dq_ws" in C.lib().x2k_last_error()


@pytest.mark.parametrize("M,N", [(300, 30522), (77, 1000), (1536, 30522)])
def test_fused_vocab_cross_entropy(dev, M, N):
    """Vocabulary GEMM fused with the online-softmax cross entropy (x2k_gemm ce_mode 1/2 + x2k_ce_finalize; models/xbert.py
    :805-834,1653-1661) against F.cross_entropy on fp32 logits of the same bf16 operands: per-row losses (ignored rows 0),
    and the gradients w.r.t. the hidden states, the tied decoder weight and the bias — no [M, vocab] logits in between."""
    import torch.nn.functional as F
    from x2vlm_b200 import functional as XF
    from x2vlm_b200.params import Shadow
    g = torch.Generator(device=dev).manual_seed(5)
    K = 768
    w = torch.nn.Parameter(_bf(torch.randn(N, K, device=dev, generator=g) * 0.05).float())
    bias = torch.nn.Parameter(torch.randn(N, device=dev, generator=g) * 0.1)
    h = _bf(torch.randn(M, K, device=dev, generator=g)).float().requires_grad_(True)
    labels = torch.randint(0, N, (M,), device=dev, generator=g)
    labels[::7] = -100
    labels[1] = N - 1          # label inside the ragged last 16-column group
    labels[2] = 0
    up = torch.rand(M, device=dev, generator=g) / M
    rows = XF.vocab_cross_entropy(h, Shadow(w), bias, labels)
    (rows * up).sum().backward()
    got = (rows.detach(), h.grad.clone(), w.grad.clone(), bias.grad.clone())
    h.grad = w.grad = bias.grad = None
    logits = h @ w.t() + bias
    ref = F.cross_entropy(logits, labels, reduction="none", ignore_index=-100)
    (ref * up).sum().backward()
    assert (got[0] - ref).abs().max().item() < 2e-3, (got[0] - ref).abs().max().item()
    assert float(got[0][::7].abs().max()) == 0.0
    for name, a, b in (("dh", got[1], h.grad), ("dW", got[2], w.grad), ("dbias", got[3], bias.grad)):
        err = ((a - b).norm() / b.norm()).item()
        assert err < 1e-2, (name, err)


@pytest.mark.parametrize("D,p_drop", [(768, 0.0), (1024, 0.1), (128, 0.0)])
def test_embed_ln_fused(dev, D, p_drop):
    """x2k_embed_ln_{fwd,bwd} against word + position + token-type gathers -> LayerNorm(1e-12) -> dropout in torch
    (models/xbert.py:189-216), the dropout mask taken from the Philox oracle."""
    import torch.nn.functional as F
    from oracle import philox
    from x2vlm_b200 import functional as XF
    g = torch.Generator(device=dev).manual_seed(3)
    B, L, V, P, T = 5, 37, 1000, 64, 2
    word = torch.nn.Parameter(torch.randn(V, D, device=dev, generator=g))
    pos = torch.nn.Parameter(torch.randn(P, D, device=dev, generator=g))
    typ = torch.nn.Parameter(torch.randn(T, D, device=dev, generator=g))
    lw = torch.nn.Parameter(1 + 0.1 * torch.randn(D, device=dev, generator=g))
    lb = torch.nn.Parameter(0.1 * torch.randn(D, device=dev, generator=g))
    ids = torch.randint(0, V, (B, L), device=dev, generator=g)
    ids[0, :5] = ids[1, :5]                      # repeated ids: the word-table scatter must accumulate
    tt = torch.randint(0, T, (B, L), device=dev, generator=g)
    dy = torch.randn(B, L, D, device=dev, generator=g)
    XF.manual_seed(77)
    y = XF.embed_ln(ids, tt, None, 3, word, pos, typ, lw, lb, 1e-12, p_drop, True)
    y.backward(dy)
    got = [y.detach()] + [q.grad.clone() for q in (word, pos, typ, lw, lb)]
    for q in (word, pos, typ, lw, lb):
        q.grad = None
    keep = 1.0
    if p_drop > 0:
        keep = torch.from_numpy(philox.keep_scale(77, 0, B * L * D, p_drop)).view(B, L, D).to(dev)
    x = word[ids] + pos[3 + torch.arange(L, device=dev)][None] + typ[tt]
    ref = F.layer_norm(x, (D,), lw, lb, 1e-12) * keep
    ref.backward(dy)
    want = [ref.detach()] + [q.grad for q in (word, pos, typ, lw, lb)]
    for name, a, b in zip(("y", "dword", "dpos", "dtype", "dln_w", "dln_b"), got, want):
        err = ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
        assert err < 2e-5, (name, err)


@pytest.mark.parametrize("D", [768, 1024])
def test_pool_tail_fused(dev, D):
    """x2k_pool_tail_{fwd,bwd} against drop-cls -> fc_norm -> mean (models/beit2.py:409-425) and the region-mode gather +
    mask-weighted mean (:430-434), forward and backward (shared images accumulate)."""
    import torch.nn.functional as F
    from x2vlm_b200 import functional as XF, synth
    g = torch.Generator(device=dev).manual_seed(9)
    n_img, N = 3, 197
    x = torch.randn(n_img, N, D, device=dev, generator=g).requires_grad_(True)
    w = torch.nn.Parameter(1 + 0.1 * torch.randn(D, device=dev, generator=g))
    b = torch.nn.Parameter(0.1 * torch.randn(D, device=dev, generator=g))
    rb = synth.region_batch(n_img, 7, 24, seed=2)
    group, atts = rb["idx_to_group_img"].to(dev), rb["image_atts"].to(dev)
    d_full = torch.randn(n_img, N, D, device=dev, generator=g)
    d_reg = torch.randn(7, N, D, device=dev, generator=g)
    full = XF.pool_tail(x, w, b, 1e-6)
    reg = XF.pool_tail(x, w, b, 1e-6, group, atts)
    ((full * d_full).sum() + (reg * d_reg).sum()).backward()
    got = [full.detach(), reg.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone()]
    x.grad = w.grad = b.grad = None
    xn = F.layer_norm(x[:, 1:], (D,), w, b, 1e-6)
    full_r = torch.cat([xn.mean(1, keepdim=True), xn], 1)
    xb = xn[group]
    wt = atts[:, 1:].unsqueeze(2).float()
    reg_r = torch.cat([(wt * xb).sum(1, keepdim=True) / wt.sum(1, keepdim=True), xb], 1)
    ((full_r * d_full).sum() + (reg_r * d_reg).sum()).backward()
    want = [full_r.detach(), reg_r.detach(), x.grad, w.grad, b.grad]
    for name, a, b_ in zip(("full", "region", "dx", "dw", "db"), got, want):
        err = ((a - b_).norm() / b_.norm().clamp_min(1e-20)).item()
        assert err < 2e-5, (name, err)
    assert float(got[2][:, 0].abs().max()) == 0.0   # the cls output is dropped: no gradient reaches it


@pytest.mark.parametrize("M,N,K", [(394, 768, 768), (2307, 1024, 256), (5000, 768, 3072)])
def test_gemm_tma_residual_epilogues(dev, M, N, K):
    """The five fp32 residual-stream epilogues of the CTA-pair kernel (TMA-loaded residual tiles, TMA-stored output tiles):
    ragged M (partial last tiles are clipped by the tensor maps), several tiles per CTA (the prefetch chain crosses tile
    boundaries), strided residual / output views."""
    from x2vlm_b200 import ops
    from oracle import philox
    g = torch.Generator(device=dev).manual_seed(1)
    A = _bf(torch.randn(M, K, device=dev, generator=g)); W = _bf(torch.randn(N, K, device=dev, generator=g) * 0.05)
    bias = torch.randn(N, device=dev, generator=g); gamma = torch.randn(N, device=dev, generator=g)
    res_buf = torch.randn(M, 2 * N, device=dev, generator=g)
    res = res_buf[:, N:]                                    # strided view: ld = 2N
    rs = torch.rand((M + 196) // 197, device=dev, generator=g) + 0.5
    acc = A.float() @ W.float().t()
    tol = 2e-4 * max(1.0, float(acc.abs().max()))
    out_buf = torch.full((M + 3, N + 64), float("nan"), device=dev)
    out = out_buf[:M, :N]                                   # the kernel must not write outside [M, N]
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    rsr = rs.repeat_interleave(197)[:M, None]
    # variant 3: bias + pre-activation + LayerScale / DropPath + residual -> fp32 (BEiT proj / fc2)
    ops.gemm(A, W, M, N, K, bias=bias, preact_out=pre, gamma=gamma, row_scale=rs, rows_per_scale=197, residual=res, out_f32=out)
    assert (out - (res + (acc + bias) * gamma * rsr)).abs().max() < tol
    assert (pre.float() - (acc + bias)).abs().max() < 0.04 * max(1.0, float(acc.abs().max()) / 8)
    assert torch.isnan(out_buf[M:]).all() and torch.isnan(out_buf[:, N:]).all()
    # variant 4: the same without LayerScale / DropPath
    out.fill_(float("nan")); pre.zero_()
    ops.gemm(A, W, M, N, K, bias=bias, preact_out=pre, residual=res, out_f32=out)
    assert (out - (res + acc + bias)).abs().max() < tol and (pre.float() - (acc + bias)).abs().max() < 0.04 * max(1.0, float(acc.abs().max()) / 8)
    # variant 5: bias + dropout + residual (BERT dense), exact Philox mask
    out.fill_(float("nan"))
    ops.gemm(A, W, M, N, K, bias=bias, dropout_p=0.1, dropout_seed=99, dropout_offset=5, residual=res, out_f32=out)
    keep = torch.from_numpy(philox.keep_scale(99, 5, M * N, 0.1)).view(M, N).to(dev)
    assert (out - (res + (acc + bias) * keep)).abs().max() < tol
    # variant 6: bias + residual (eval), variant 9: residual only (dgrad + residual-stream gradient)
    out.fill_(float("nan"))
    ops.gemm(A, W, M, N, K, bias=bias, residual=res, out_f32=out)
    assert (out - (res + acc + bias)).abs().max() < tol
    out.fill_(float("nan"))
    ops.gemm(A, W, M, N, K, residual=res, out_f32=out)
    assert (out - (res + acc)).abs().max() < tol
    assert torch.isnan(out_buf[M:]).all() and torch.isnan(out_buf[:, N:]).all()
