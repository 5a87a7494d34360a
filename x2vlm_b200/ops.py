"""Thin Python wrappers over the C ABI (include/x2k.h).  Each wrapper validates torch tensors, hands
raw device pointers + sizes to libx2k.so and launches on torch's current CUDA stream.  There is no
fallback: CPU tensors or a missing library raise.
"""
import ctypes

import torch

from . import _capi as C

_VP = ctypes.c_void_p


def _stream():
    return _VP(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return _VP(t.data_ptr()) if t is not None else None


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise C.X2kError("%s must be a CUDA tensor (the x2k kernels have no CPU path)" % name)
    if t.dtype != dtype:
        raise C.X2kError("%s must be %s, got %s" % (name, dtype, t.dtype))


def _row_major(t, name):
    if t is not None and t.dim() >= 2 and t.stride(-1) != 1:
        raise C.X2kError("%s must have a contiguous last dimension" % name)


def launch_count():
    return int(C.lib().x2k_launch_count())


# -- optional per-launch device timing (bench.py's roofline leg) -------------------------------------------------
_timing = None  # list of (family, algorithmic flops, start event, end event) while kernel_timing() is active


class kernel_timing:
    """Context manager: bracket every GEMM / attention launch with CUDA events on the launching stream.
    `totals()` afterwards gives {family: (launches, flops, ms)}.  Off (and free) by default."""

    def __enter__(self):
        global _timing
        _timing = self.records = []
        return self

    def __exit__(self, *exc):
        global _timing
        _timing = None

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for fam, flops, st, en, _tag in self.records:
            n, f, ms = out.get(fam, (0, 0.0, 0.0))
            out[fam] = (n + 1, f + flops, ms + st.elapsed_time(en))
        return out

    def by_shape(self):
        """{(family, shape tag): (launches, flops, ms)} — which shapes / epilogues run below the family average."""
        torch.cuda.synchronize()
        out = {}
        for fam, flops, st, en, tag in self.records:
            n, f, ms = out.get((fam, tag), (0, 0.0, 0.0))
            out[(fam, tag)] = (n + 1, f + flops, ms + st.elapsed_time(en))
        return out


def _timed(family, flops, call, tag=""):
    if _timing is None:
        return call()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    r = call()
    en.record()
    _timing.append((family, flops, st, en, tag() if callable(tag) else tag))
    return r


_DROP_BASE = {}


def dropout_base(device):
    """Device-side counter added to every kernel's (host-baked) dropout offset.  Zero in eager execution — the host
    counter of functional.dropout_state advances instead; a captured step graph advances it on the device at the end
    of every replay (accelerator.GraphedStep), so replays draw fresh Philox ranges although their kernel arguments are
    frozen."""
    key = (device.type, device.index)
    t = _DROP_BASE.get(key)
    if t is None:
        t = _DROP_BASE[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return t


def gemm(a, b, M, N, K, a_mn=False, b_mn=False, bias=None, act=C.ACT_NONE, aux=None, preact_out=None,
         dropout_p=0.0, dropout_seed=0, dropout_offset=0, gamma=None, row_scale=None, rows_per_scale=0,
         residual=None, accumulate=False, out_bf16=None, out_f32=None, tile_n=0, split_k=0, ce=None):
    """C[M,N] = epilogue(A·Bᵀ) — see x2k_gemm in include/x2k.h.  `a`/`b` are 2-D bf16 views whose
    stride(0) is the leading dimension; MN-major operands are passed as their stored [K, M|N] matrix."""
    _req(a, torch.bfloat16, "a"); _req(b, torch.bfloat16, "b")
    _row_major(a, "a"); _row_major(b, "b")
    for t, n in ((bias, "bias"), (gamma, "gamma"), (row_scale, "row_scale"), (residual, "residual"), (out_f32, "out_f32")):
        _req(t, torch.float32, n)
    for t, n in ((aux, "aux"), (preact_out, "preact_out"), (out_bf16, "out_bf16")):
        _req(t, torch.bfloat16, n)
    g = C.X2kGemmArgs()
    g.A, g.B = a.data_ptr(), b.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb = a.stride(0), b.stride(0)
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    if bias is not None:
        g.bias = bias.data_ptr()
    g.act = act
    if aux is not None:
        g.aux, g.ld_aux = aux.data_ptr(), aux.stride(0)
    if preact_out is not None:
        g.preact_out, g.ld_preact = preact_out.data_ptr(), preact_out.stride(0)
    g.dropout_p, g.dropout_seed, g.dropout_offset = float(dropout_p), int(dropout_seed), int(dropout_offset)
    if dropout_p > 0.0:
        g.dropout_offset_dev = dropout_base(a.device).data_ptr()
    if gamma is not None:
        g.gamma = gamma.data_ptr()
    if row_scale is not None:
        g.row_scale, g.rows_per_scale = row_scale.data_ptr(), int(rows_per_scale)
    if residual is not None:
        g.residual, g.ld_res = residual.data_ptr(), residual.stride(0)
    g.accumulate = int(accumulate)
    if out_bf16 is not None:
        g.out_bf16, g.ld_out_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ld_out_f32 = out_f32.data_ptr(), out_f32.stride(0)
    g.tile_n = tile_n
    g.split_k = split_k
    if ce is not None:  # fused cross entropy over N: dict(mode, labels, partials, target_logit) or dict(mode, labels, lse, row_grad)
        _req(ce["labels"], torch.int64, "ce.labels")
        g.ce_mode, g.ce_labels = int(ce["mode"]), ce["labels"].data_ptr()
        for k in ("partials", "target_logit", "lse", "row_grad"):
            if ce.get(k) is not None:
                _req(ce[k], torch.float32, "ce." + k)
                setattr(g, "ce_" + k, ce[k].data_ptr())
    def tag():
        epi = "+".join(n for n, on in (("bias", bias is not None), ("preact", preact_out is not None), ("gelu", act == C.ACT_GELU),
                                       ("gelu'", act == C.ACT_GELU_BWD), ("drop", dropout_p > 0.0),
                                       ("scale", gamma is not None or row_scale is not None), ("res", residual is not None),
                                       ("acc", accumulate), ("f32", out_f32 is not None), ("bf16", out_bf16 is not None)) if on)
        return "%dx%dx%d %s%s %s" % (M, N, K, "T" if a_mn else "N", "T" if b_mn else "N", epi)
    _timed("gemm", 2.0 * M * N * K, lambda: C.check(C.lib().x2k_gemm(ctypes.byref(g), _stream()), "x2k_gemm"), tag)


def embed_ln_fwd(ids, type_ids, pos_ids, L, pos_offset, word, pos, typ, ln_w, ln_b, eps, y_f32, y_bf16, mean, rstd, dropout_p=0.0,
                 dropout_seed=0, dropout_offset=0):
    """ids / type_ids / pos_ids: flat int64 [M] (type_ids / pos_ids may be None)."""
    _req(ids, torch.int64, "ids"); _req(type_ids, torch.int64, "type_ids"); _req(pos_ids, torch.int64, "pos_ids")
    for t, n in ((word, "word"), (pos, "pos"), (typ, "type"), (ln_w, "ln_w"), (ln_b, "ln_b"), (y_f32, "y_f32")):
        _req(t, torch.float32, n)
    M, D = ids.numel(), word.shape[1]
    C.check(C.lib().x2k_embed_ln_fwd(_p(ids), _p(type_ids), _p(pos_ids), M, D, int(L), int(pos_offset), _p(word), _p(pos), _p(typ),
                                     _p(ln_w), _p(ln_b), float(eps), float(dropout_p), int(dropout_seed), int(dropout_offset),
                                     _p(dropout_base(word.device)) if dropout_p > 0.0 else None, _p(y_f32), _p(y_bf16), _p(mean),
                                     _p(rstd), _stream()), "x2k_embed_ln_fwd")


def embed_ln_bwd(dy, ids, type_ids, pos_ids, L, pos_offset, word, pos, typ, ln_w, ln_b, mean, rstd, dword, dpos, dtype, dw, db,
                 dropout_p=0.0, dropout_seed=0, dropout_offset=0):
    _req(dy, torch.float32, "dy")
    M, D = ids.numel(), word.shape[1]
    C.check(C.lib().x2k_embed_ln_bwd(_p(dy), _p(ids), _p(type_ids), _p(pos_ids), M, D, int(L), int(pos_offset), _p(word), _p(pos),
                                     _p(typ), _p(ln_w), _p(ln_b), _p(mean), _p(rstd), float(dropout_p), int(dropout_seed),
                                     int(dropout_offset), _p(dropout_base(word.device)) if dropout_p > 0.0 else None, _p(dword),
                                     _p(dpos), _p(dtype), _p(dw), _p(db), _stream()), "x2k_embed_ln_bwd")


def pool_tail_fwd(x, n_out, group, atts, w, b, eps, out, mean=None, rstd=None):
    """x fp32 [n_img, N, D] contiguous; group int64 [n_out] or None; atts int64 [n_out, N] or None."""
    _req(x, torch.float32, "x"); _req(group, torch.int64, "group"); _req(atts, torch.int64, "atts")
    n_img, N, D = x.shape
    C.check(C.lib().x2k_pool_tail_fwd(_p(x), n_img, int(n_out), N, D, _p(group), _p(atts), _p(w), _p(b), float(eps), _p(out), _p(mean),
                                      _p(rstd), _stream()), "x2k_pool_tail_fwd")


def pool_tail_bwd(d_out, x, n_out, group, atts, w, mean, rstd, dx, dw, db, accumulate=False):
    _req(d_out, torch.float32, "d_out"); _req(dx, torch.float32, "dx")
    n_img, N, D = x.shape
    C.check(C.lib().x2k_pool_tail_bwd(_p(d_out), _p(x), n_img, int(n_out), N, D, _p(group), _p(atts), _p(w), _p(mean), _p(rstd),
                                      _p(dx), int(accumulate), _p(dw), _p(db), _stream()), "x2k_pool_tail_bwd")


def ce_finalize(partials, target_logit, labels, M, N, lse, loss):
    C.check(C.lib().x2k_ce_finalize(_p(partials), _p(target_logit), _p(labels), M, N, _p(lse), _p(loss), _stream()), "x2k_ce_finalize")


def layernorm_fwd(x, w, b, eps, y_bf16=None, y_f32=None, mean=None, rstd=None):
    _req(x, torch.float32, "x"); _req(w, torch.float32, "w"); _req(b, torch.float32, "b")
    M, D = x.shape
    C.check(C.lib().x2k_layernorm_fwd(_p(x), _p(w), _p(b), M, D, float(eps), _p(y_bf16), _p(y_f32), _p(mean), _p(rstd),
                                      _stream()), "x2k_layernorm_fwd")


def layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, dx_residual=None, g_bf16=None, dbias=None, dropout_p=0.0, dropout_seed=0,
                  dropout_offset=0):
    """dy: bf16 or fp32 [M,D], or a (fp32, bf16) pair that is summed; dw/db are accumulated into.
    With g_bf16: also g = bf16(dx * dropout keep-scale) and dbias += colsum(dx * keep-scale) in the same pass
    (x2k_layernorm_bwd_dropcast)."""
    M, D = x.shape
    if isinstance(dy, (tuple, list)):
        dy_f, dy_b = dy
    else:
        dy_b = dy if dy.dtype == torch.bfloat16 else None
        dy_f = dy if dy.dtype == torch.float32 else None
    if g_bf16 is not None:
        _req(g_bf16, torch.bfloat16, "g_bf16"); _req(dbias, torch.float32, "dbias")
        if g_bf16.stride(0) != D:
            raise C.X2kError("g_bf16 must be contiguous [M, D]")
        C.check(C.lib().x2k_layernorm_bwd_dropcast(
            _p(dy_b), _p(dy_f), _p(x), _p(w), _p(mean), _p(rstd), _p(dx_residual), M, D, _p(dx), _p(dw), _p(db),
            float(dropout_p), int(dropout_seed), int(dropout_offset),
            _p(dropout_base(x.device)) if dropout_p > 0.0 else None, _p(g_bf16), _p(dbias), _stream()),
            "x2k_layernorm_bwd_dropcast")
        return
    C.check(C.lib().x2k_layernorm_bwd(_p(dy_b), _p(dy_f), _p(x), _p(w), _p(mean), _p(rstd), _p(dx_residual), M, D, _p(dx),
                                      _p(dw), _p(db), _stream()), "x2k_layernorm_bwd")


def scale_cast_colsum(dx, M, N, g_bf16=None, gamma=None, row_scale=None, rows_per_scale=0, dropout_p=0.0, dropout_seed=0,
                      dropout_offset=0, y_bf16=None, dbias=None, dgamma=None):
    _req(dx, torch.float32, "dx")
    C.check(C.lib().x2k_scale_cast_colsum(
        _p(dx), dx.stride(0), M, N, _p(gamma), _p(row_scale), int(rows_per_scale), float(dropout_p), int(dropout_seed),
        int(dropout_offset), _p(dropout_base(dx.device)) if dropout_p > 0.0 else None, _p(y_bf16), y_bf16.stride(0) if y_bf16 is not None else 0, _p(g_bf16),
        g_bf16.stride(0) if g_bf16 is not None else 0, _p(dbias), _p(dgamma), _stream()), "x2k_scale_cast_colsum")


def colsum_bf16(x, M, N, dcol):
    _req(x, torch.bfloat16, "x"); _req(dcol, torch.float32, "dcol")
    C.check(C.lib().x2k_colsum_bf16(_p(x), x.stride(0), M, N, _p(dcol), _stream()), "x2k_colsum_bf16")


def segment_sum_bf16(src, index, n_seg, out):
    """out[s] = sum of src rows b with index[b] == s; src [n_rows, row_elems] bf16 contiguous."""
    _req(src, torch.bfloat16, "src"); _req(index, torch.int32, "index"); _req(out, torch.bfloat16, "out")
    n_rows = src.shape[0]
    C.check(C.lib().x2k_segment_sum_bf16(_p(src), _p(index), n_rows, src.numel() // n_rows, int(n_seg), _p(out), _stream()),
            "x2k_segment_sum_bf16")


def cast_f32_bf16(src, dst, n=None):
    _req(src, torch.float32, "src"); _req(dst, torch.bfloat16, "dst")
    C.check(C.lib().x2k_cast_f32_bf16(_p(src), _p(dst), int(n if n is not None else src.numel()), _stream()),
            "x2k_cast_f32_bf16")


def to_bf16(x):
    """fp32 -> bf16 copy through the x2k cast kernel (contiguous tensors)."""
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.numel():
        cast_f32_bf16(x, out)
    return out


def attn_group_table(kv_index, n_kv, Lq, Lk):
    """Work-item table of the grouped cross-attention kernels (x2k_attn_group_build): the B query sequences
    grouped by the K/V source kv_index[b] they attend to.  Returns None when (Lq, Lk) is not eligible for the
    grouped path (the one-sequence-per-tile kernels are used then)."""
    _req(kv_index, torch.int32, "kv_index")
    B = kv_index.numel()
    if C.lib().x2k_attn_group_slots(Lq) < 1 or pad16(Lk) > 256 or n_kv > 4096:
        return None
    n = C.lib().x2k_attn_group_table_ints(B, n_kv, Lq)
    table = torch.empty(n, dtype=torch.int32, device=kv_index.device)
    C.check(C.lib().x2k_attn_group_build(_p(kv_index), B, n_kv, Lq, _p(table), _stream()), "x2k_attn_group_build")
    return table


def _attn_args(q, k, v, B, H, Lq, Lk, scale, o, lse, kv_index=None, n_kv=0, bias=None, mask=None, mask_per_query=False,
               dropout_p=0.0, dropout_seed=0, dropout_offset=0, kv_groups=None):
    a = C.X2kAttnArgs()
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o")):
        _req(t, torch.bfloat16, n)
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ld_q, a.ld_k, a.ld_v = q.stride(0), k.stride(0), v.stride(0)
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    if kv_index is not None:
        _req(kv_index, torch.int32, "kv_index")
        a.kv_index = kv_index.data_ptr()
    if kv_groups is not None:
        _req(kv_groups, torch.int32, "kv_groups")
        a.kv_groups = kv_groups.data_ptr()
    a.n_kv = int(n_kv)
    a.scale = float(scale)
    if bias is not None:  # [H, Lq, ld]
        _req(bias, torch.float32, "bias")
        a.bias, a.bias_h_stride, a.bias_q_stride = bias.data_ptr(), bias.stride(0), bias.stride(1)
    if mask is not None:  # [B, ld] or [B, Lq, ld]
        _req(mask, torch.float32, "mask")
        a.mask, a.mask_b_stride = mask.data_ptr(), mask.stride(0)
        a.mask_q_stride = mask.stride(1) if mask_per_query else 0
    a.dropout_p, a.dropout_seed, a.dropout_offset = float(dropout_p), int(dropout_seed), int(dropout_offset)
    if dropout_p > 0.0:
        a.dropout_offset_dev = dropout_base(q.device).data_ptr()
    a.o, a.ld_o = o.data_ptr(), o.stride(0)
    a.lse = lse.data_ptr()
    return a


def attn_fwd(q, k, v, B, H, Lq, Lk, scale, o, lse, **kw):
    """q/k/v/o: 2-D bf16 views [rows, >= H*64] (rows = B*Lq or n_kv*Lk) with row stride = ld."""
    a = _attn_args(q, k, v, B, H, Lq, Lk, scale, o, lse, **kw)
    _timed("attn_fwd", 4.0 * B * H * Lq * Lk * 64,
           lambda: C.check(C.lib().x2k_attn_fwd(ctypes.byref(a), _stream()), "x2k_attn_fwd"), "B%d Lq%d Lk%d" % (B, Lq, Lk))


def attn_bwd(q, k, v, B, H, Lq, Lk, scale, o, lse, d_o, dq, dk, dv, ds_out=None, **kw):
    """dk/dv: per query sequence [B*Lk rows]; with kv_groups=... per K/V source [n_kv*Lk rows], already summed."""
    a = _attn_args(q, k, v, B, H, Lq, Lk, scale, o, lse, **kw)
    for t, n in ((d_o, "d_o"), (dq, "dq"), (dk, "dk"), (dv, "dv"), (ds_out, "ds_out")):
        _req(t, torch.bfloat16, n)
    a.d_o, a.ld_do = d_o.data_ptr(), d_o.stride(0)
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.ld_dq, a.ld_dk, a.ld_dv = dq.stride(0), dk.stride(0), dv.stride(0)
    delta_ws = torch.empty(B, H, Lq, dtype=torch.float32, device=q.device)  # rowsum(dO * O), filled by a pre-kernel
    a.delta_ws = delta_ws.data_ptr()
    ws_bytes = int(C.lib().x2k_attn_bwd_workspace_bytes(ctypes.byref(a)))
    if ws_bytes:  # key-blocked kernels (Lq or Lk > 256): fp32 dQ accumulator shared by the key-block CTAs
        dq_ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=q.device)
        a.dq_ws = dq_ws.data_ptr()
    if ds_out is not None:  # [B, H, Lq, ld]
        a.ds_out = ds_out.data_ptr()
        a.ds_b_stride, a.ds_h_stride, a.ds_q_stride = ds_out.stride(0), ds_out.stride(1), ds_out.stride(2)
    _timed("attn_bwd", 10.0 * B * H * Lq * Lk * 64,
           lambda: C.check(C.lib().x2k_attn_bwd(ctypes.byref(a), _stream()), "x2k_attn_bwd"), "B%d Lq%d Lk%d" % (B, Lq, Lk))


def attn_probs(q, k, B, H, Lq, Lk, scale, lse, **kw):
    """fp32 [B, H, Lq, Lk]: the pre-dropout attention probabilities rebuilt from Q, K and the forward's lse."""
    o_dummy = q  # not read in this mode
    a = _attn_args(q, k, k, B, H, Lq, Lk, scale, o_dummy, lse, **kw)
    out = torch.empty(B, H, Lq, Lk, dtype=torch.float32, device=q.device)
    C.check(C.lib().x2k_attn_probs(ctypes.byref(a), 0, _p(out), _stream()), "x2k_attn_probs")
    return out


def attn_probs_grad(d_o, v, B, H, Lq, Lk, **kw):
    """fp32 [B, H, Lq, Lk]: dL/d(attention probabilities) = (dO · Vᵀ) * dropout keep-scale."""
    a = _attn_args(d_o, v, v, B, H, Lq, Lk, 1.0, d_o, torch.empty(0, device=d_o.device), **kw)
    a.d_o, a.ld_do = d_o.data_ptr(), d_o.stride(0)
    out = torch.empty(B, H, Lq, Lk, dtype=torch.float32, device=d_o.device)
    C.check(C.lib().x2k_attn_probs(ctypes.byref(a), 1, _p(out), _stream()), "x2k_attn_probs")
    return out


def relpos_bias_gather(table, index, N, H, out):
    _req(table, torch.float32, "table"); _req(index, torch.int64, "index"); _req(out, torch.float32, "out")
    C.check(C.lib().x2k_relpos_bias_gather(_p(table), _p(index), N, H, _p(out), out.stride(1), _stream()),
            "x2k_relpos_bias_gather")


def relpos_bias_scatter(ds, B, H, N, index, dtable):
    _req(ds, torch.bfloat16, "ds"); _req(dtable, torch.float32, "dtable")
    C.check(C.lib().x2k_relpos_bias_scatter(_p(ds), B, H, N, ds.stride(0), ds.stride(1), ds.stride(2), _p(index), _p(dtable),
                                            _stream()), "x2k_relpos_bias_scatter")


def sumsq(g, out, n=None):
    C.check(C.lib().x2k_sumsq(_p(g), int(n if n is not None else g.numel()), _p(out), _stream()), "x2k_sumsq")


def adamw_flat(p, g, m, v, p_bf16, n, seg_end, seg_lr, seg_wd, beta1, beta2, eps, step=0, step_dev=None, grad_scale=None,
               zero_grad=False):
    C.check(C.lib().x2k_adamw_flat(_p(p), _p(g), _p(m), _p(v), _p(p_bf16), int(n), _p(seg_end), _p(seg_lr), _p(seg_wd),
                                   int(seg_end.numel()), float(beta1), float(beta2), float(eps), int(step), _p(step_dev),
                                   _p(grad_scale), int(zero_grad), _stream()), "x2k_adamw_flat")


def pad16(n):
    return (n + 15) // 16 * 16


def pad32(n):
    """Row stride (in elements) of attention bias / mask buffers: the kernels read them in 32-column units."""
    return (n + 31) // 32 * 32
