"""Autograd functions of the hot path, each a fixed sequence of libx2k kernel launches.

One function per transformer block (BEiT block, BERT text/fusion layer) plus a generic fused Linear:
forward and backward are written out by hand (dgrad / wgrad GEMMs read activations and weights where
they lie through K-major / MN-major UMMA descriptors; LayerNorm, GELU', dropout, LayerScale, DropPath
and residual adds live in kernel prologues/epilogues), so autograd sees ONE node per block instead of
the reference's ~34 (BEiT) / ~56 (fusion) ATen ops per block.

Semantics restated from models/beit2.py:125-209 and models/xbert.py:322-625; parity is checked in
tests/ against oracle/restate.py (itself pinned against the unmodified reference).
"""
import math
import threading

import torch

from . import ops
from ._capi import ACT_GELU, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, ACT_NONE


# ----------------------------------------------------------------------------------------------
# dropout bookkeeping: every dropout site gets a unique Philox (seed, offset) range
# ----------------------------------------------------------------------------------------------
class _DropoutState:
    """Philox key + running offset of the in-kernel dropout masks.  Unless `manual_seed` pins it, the key is derived on
    first use from torch's seed and the data-parallel rank — the reference flow `torch.manual_seed(args.seed + rank)`
    (Pretrain.py:437-441) therefore gives every rank, and every --seed, its own masks."""

    def __init__(self):
        self.seed = None
        self.offset = 0
        self.lock = threading.Lock()

    def manual_seed(self, seed):
        self.seed, self.offset = int(seed) & 0xFFFFFFFFFFFFFFFF, 0

    def _default_seed(self):
        import os
        rank = int(os.environ.get("RANK", "0"))
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank = dist.get_rank()
        except Exception:
            pass
        return ((torch.initial_seed() * 0x9E3779B97F4A7C15) ^ ((rank + 1) * 0xD1B54A32D192ED03) ^ 0x5EED5EED) & 0xFFFFFFFFFFFFFFFF

    def take(self, n_elems):
        with self.lock:
            if self.seed is None:
                self.seed = self._default_seed()
            off = self.offset
            self.offset += (int(n_elems) + 3) // 4 + 1
        return self.seed, off


dropout_state = _DropoutState()


def manual_seed(seed):
    """Seed the Philox stream of the in-kernel dropout masks."""
    dropout_state.manual_seed(seed)


def _zeros(n, dev):
    return torch.zeros(n, dtype=torch.float32, device=dev)


def _empty_bf16(*shape, dev):
    return torch.empty(*shape, dtype=torch.bfloat16, device=dev)


class _G:
    """Gradient slot of a small parameter (bias / LayerNorm / LayerScale / rel-pos table) inside a fused backward:
    the arena's flat-gradient view — the kernel accumulates in place and autograd receives None — or, without an
    arena, a fresh zero buffer that is returned to autograd."""
    __slots__ = ("param", "buf", "direct", "frozen")

    def __init__(self, param, dev):
        self.param = param
        self.frozen = param is not None and not param.requires_grad
        g = getattr(param, "_x2k_grad", None) if param is not None else None
        if g is not None and not self.frozen:
            self.buf, self.direct = g.view(-1), True
        else:  # frozen parameters (fine-tuning with frozen layers): the kernels still need a target, nothing keeps it
            self.direct = False
            self.buf = torch.zeros(param.numel(), dtype=torch.float32, device=dev) if param is not None else None

    def ret(self):
        if self.param is None or self.frozen:
            return None
        if self.direct:
            self.param._x2k_arena.note_grad_written(self.param, [self.param])
            return None
        return self.buf.view(self.param.shape)


def _note_uses(*objs):
    """Count one forward use of each shadow / arena-managed parameter (only while autograd is recording), so the DDP
    bucketer knows when the LAST gradient contribution of a step has been written."""
    if not torch.is_grad_enabled():
        return
    for o in objs:
        if o is None:
            continue
        a = getattr(o, "arena", None) if not isinstance(o, torch.Tensor) else getattr(o, "_x2k_arena", None)
        if a is not None:
            a.note_use(o)


def _wgrad(shadow, dy_bf16, x_bf16, n_out, n_in, rows):
    """dW[n_out, n_in] = dyᵀ · x over `rows`; into the arena's flat gradient (accumulating) or a new tensor.
    Returns the list of per-parameter grads to hand to autograd (None when the sink took them)."""
    if not any(q.requires_grad for q in shadow.params):  # frozen weights: no wgrad GEMM, nothing enters the clip norm
        return [None] * len(shadow.params)
    sink = shadow.grad_sink()
    if sink is not None:
        ops.gemm(dy_bf16, x_bf16, n_out, n_in, rows, a_mn=True, b_mn=True, out_f32=sink, accumulate=True)
        shadow.grads_done()
        return [None] * len(shadow.params)
    g = torch.empty(n_out, n_in, dtype=torch.float32, device=dy_bf16.device)
    ops.gemm(dy_bf16, x_bf16, n_out, n_in, rows, a_mn=True, b_mn=True, out_f32=g)
    return shadow.split_grad(g)


# ----------------------------------------------------------------------------------------------
# generic fused Linear:  y = act(x · Wᵀ + b)
# ----------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, shadow, gelu, out_bf16, pad_value, *weights):
        # x: [M, K] fp32 or bf16; weights only listed so autograd tracks them.  pad_value (float or None): when the
        # output width N is not a multiple of 8 the kernel writes rows of ld = pad8(N) elements; with pad_value the
        # PADDED [M, ld] tensor is returned, its extra columns filled (e.g. -inf for logits that go straight into a
        # softmax / cross entropy) — no compaction copy forward, no re-padding copy backward.
        dev = x.device
        M, K = x.shape
        N = shadow.total_rows
        xb = x if x.dtype == torch.bfloat16 else ops.to_bf16(x)
        if xb.stride(-1) != 1 or xb.stride(0) % 8 != 0:
            xb = xb.contiguous()
        w = shadow.get()
        ld = (N + 7) // 8 * 8
        pre = _empty_bf16(M, ld, dev=dev) if gelu else None
        if out_bf16:
            y = _empty_bf16(M, ld, dev=dev)
            ops.gemm(xb, w, M, N, K, bias=bias, act=ACT_GELU_SAVE_GRAD if gelu else ACT_NONE, preact_out=pre, out_bf16=y)
        else:
            y = torch.empty(M, ld, dtype=torch.float32, device=dev)
            ops.gemm(xb, w, M, N, K, bias=bias, act=ACT_GELU_SAVE_GRAD if gelu else ACT_NONE, preact_out=pre, out_f32=y)
        ctx.shadow, ctx.gelu, ctx.dims, ctx.x_dtype = shadow, gelu, (M, N, K, ld), x.dtype
        ctx.has_bias = bias is not None
        ctx.padded_out = pad_value is not None and ld != N
        ctx.save_for_backward(xb, pre)
        if ctx.padded_out:
            y[:, N:].fill_(pad_value)
            return y
        return y[:, :N] if ld != N else y

    @staticmethod
    def backward(ctx, dy):
        xb, pre = ctx.saved_tensors
        M, N, K, ld = ctx.dims
        dev = dy.device
        shadow = ctx.shadow
        # dy -> bf16 [M, ld] (zero padded columns)
        if ctx.padded_out:  # dy is already [M, ld]; its pad columns carry no gradient by construction of the caller
            dyb = ops.to_bf16(dy.float() if dy.dtype != torch.float32 else dy)
            dyb[:, N:].zero_()
        elif ld != N:
            dyb = torch.zeros(M, ld, dtype=torch.bfloat16, device=dev)
            dyb[:, :N].copy_(dy)
        elif dy.dtype == torch.bfloat16 and dy.is_contiguous():
            dyb = dy
        else:
            dyb = ops.to_bf16(dy.float() if dy.dtype != torch.float32 else dy)
        if ctx.gelu:
            # dpre = dy * gelu'(pre); `pre` holds gelu'(pre-activation) itself (X2K_ACT_GELU_SAVE_GRAD).  Small uses only
            # (MLM transform), so the product is taken elementwise in torch instead of a GEMM epilogue.
            dpre = dyb[:, :N].float() * pre[:, :N].float()
            dyb = torch.zeros(M, ld, dtype=torch.bfloat16, device=dev)
            dyb[:, :N].copy_(dpre)
        dbias = None
        if ctx.has_bias:
            dbias = dyb[:, :N].float().sum(0)
        dx = None
        if ctx.needs_input_grad[0]:
            w = shadow.get_nograd()
            if ctx.x_dtype == torch.bfloat16:
                dx = _empty_bf16(M, K, dev=dev)
                ops.gemm(dyb, w, M, K, N, b_mn=True, out_bf16=dx)
            else:
                dx = torch.empty(M, K, dtype=torch.float32, device=dev)
                ops.gemm(dyb, w, M, K, N, b_mn=True, out_f32=dx)
        wg = _wgrad(shadow, dyb, xb, N, K, M)
        return (dx, dbias, None, None, None, None, *wg)


def linear(x, shadow, bias=None, gelu=False, out_bf16=False, pad_value=None):
    """y[M,N] = act(x[M,K] · Wᵀ + b) on the tcgen05 GEMM; x fp32 or bf16, W given by its Shadow.  With pad_value the
    result is [M, pad8(N)] and the extra columns hold pad_value (see _LinearFn.forward)."""
    _note_uses(shadow)
    return _LinearFn.apply(x, bias, shadow, gelu, out_bf16, pad_value, *shadow.params)


# ----------------------------------------------------------------------------------------------
# BertEmbeddings (models/xbert.py:189-216) and the vision tail (models/beit2.py:409-436) as single kernels
# ----------------------------------------------------------------------------------------------
class _EmbedLnFn(torch.autograd.Function):
    """dropout(LayerNorm(word[ids] + pos[p] + type[t])): one gather + LayerNorm + dropout kernel forward, one kernel
    backward that recomputes the row and scatters its gradient into the three tables (x2k_embed_ln_{fwd,bwd})."""

    @staticmethod
    def forward(ctx, ids, type_ids, pos_ids, pos_offset, eps, drop, word, pos, typ, ln_w, ln_b):
        B, L = ids.shape
        M, D, dev = B * L, word.shape[1], word.device
        flat = lambda t: t.reshape(-1).contiguous() if t is not None else None
        ids_f, type_f = flat(ids), flat(type_ids)
        pos_f = flat(pos_ids.expand(B, L)) if pos_ids is not None else None
        y = torch.empty(M, D, dtype=torch.float32, device=dev)
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.embed_ln_fwd(ids_f, type_f, pos_f, L, pos_offset, word, pos, typ, ln_w, ln_b, eps, y, None, mean, rstd,
                         dropout_p=drop[0], dropout_seed=drop[1], dropout_offset=drop[2])
        ctx.meta = (L, pos_offset, drop)
        ctx.P = (word, pos, typ, ln_w, ln_b)
        ctx.save_for_backward(ids_f, type_f, pos_f, mean, rstd, word, pos, typ, ln_w, ln_b)
        return y.view(B, L, D)

    @staticmethod
    def backward(ctx, dy):
        ids_f, type_f, pos_f, mean, rstd, word, pos, typ, ln_w, ln_b = ctx.saved_tensors
        L, pos_offset, drop = ctx.meta
        dev = dy.device
        G = [_G(q, dev) for q in ctx.P]
        dy2 = dy.contiguous().view(-1, dy.shape[-1])
        ops.embed_ln_bwd(dy2, ids_f, type_f, pos_f, L, pos_offset, word, pos, typ, ln_w, ln_b, mean, rstd, G[0].buf, G[1].buf,
                         G[2].buf, G[3].buf, G[4].buf, dropout_p=drop[0], dropout_seed=drop[1], dropout_offset=drop[2])
        return (None, None, None, None, None, None, *[g.ret() for g in G])


def embed_ln(ids, type_ids, pos_ids, pos_offset, word, pos, typ, ln_w, ln_b, eps, p_drop, training):
    """ids [B, L] int64 -> fp32 [B, L, D].  pos_ids None = pos_offset + arange(L) (BertEmbeddings' default)."""
    _note_uses(word, pos, typ, ln_w, ln_b)
    drop = _drop(p_drop, training, ids.numel() * word.shape[1])
    return _EmbedLnFn.apply(ids, type_ids, pos_ids, int(pos_offset), float(eps), drop, word, pos, typ, ln_w, ln_b)


class _PoolTailFn(torch.autograd.Function):
    """[n_out, N, D]: fc_norm of the patch tokens of image group[s] (cls output dropped) with their (mask-weighted) mean
    as token 0 (x2k_pool_tail_{fwd,bwd})."""

    @staticmethod
    def forward(ctx, x, group, atts, eps, w, b):
        x = x.contiguous()
        n_img, N, D = x.shape
        n_out = group.numel() if group is not None else n_img
        out = torch.empty(n_out, N, D, dtype=torch.float32, device=x.device)
        need = any(ctx.needs_input_grad)  # LayerNorm statistics are only kept for a backward pass
        mean = torch.empty(n_out, N, device=x.device) if need else None
        rstd = torch.empty(n_out, N, device=x.device) if need else None
        atts_c = atts.contiguous() if atts is not None else None
        ops.pool_tail_fwd(x, n_out, group, atts_c, w, b, eps, out, mean, rstd)
        ctx.P = (w, b)
        ctx.n_out = n_out
        ctx.save_for_backward(x, group, atts_c, w, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, group, atts, w, mean, rstd = ctx.saved_tensors
        G = [_G(q, d_out.device) for q in ctx.P]
        dx = torch.empty_like(x)
        ops.pool_tail_bwd(d_out.contiguous(), x, ctx.n_out, group, atts, w, mean, rstd, dx, G[0].buf, G[1].buf)
        return (dx, None, None, None, G[0].ret(), G[1].ret())


def pool_tail(x, w, b, eps, group=None, atts=None):
    """Vision tail on the last block's output x [n_img, N, D] (fp32): see _PoolTailFn."""
    _note_uses(w, b)
    return _PoolTailFn.apply(x, group, atts, float(eps), w, b)


# ----------------------------------------------------------------------------------------------
# vocabulary projection fused with the cross entropy
# ----------------------------------------------------------------------------------------------
class _VocabCEFn(torch.autograd.Function):
    """loss[m] = CrossEntropy(h[m] · Wᵀ + b, labels[m]) per row (0 where labels[m] < 0), without ever storing the
    [M, vocab] logits: the GEMM epilogue emits online-softmax statistics per 16-column group (x2k_gemm ce_mode 1), a
    streaming kernel reduces them (x2k_ce_finalize); backward recomputes the logits tile by tile and writes
    dlogits = g[m] · (softmax - onehot) straight as the bf16 operand of the dgrad / wgrad GEMMs (ce_mode 2).
    Replaces decoder Linear + CrossEntropyLoss of models/xbert.py:805-834,1653-1661."""

    @staticmethod
    def forward(ctx, h, bias, labels, shadow, *weights):
        dev = h.device
        M, K = h.shape
        N = shadow.total_rows
        hb = h if h.dtype == torch.bfloat16 else ops.to_bf16(h)
        labels = labels.contiguous()
        groups = (N + 15) // 16
        partials = torch.empty(M, groups, 2, dtype=torch.float32, device=dev)
        tlogit = torch.zeros(M, dtype=torch.float32, device=dev)
        ops.gemm(hb, shadow.get(), M, N, K, bias=bias, ce=dict(mode=1, labels=labels, partials=partials, target_logit=tlogit))
        lse = torch.empty(M, dtype=torch.float32, device=dev)
        loss = torch.empty(M, dtype=torch.float32, device=dev)
        ops.ce_finalize(partials, tlogit, labels, M, N, lse, loss)
        ctx.shadow, ctx.dims, ctx.h_dtype, ctx.has_bias = shadow, (M, N, K), h.dtype, bias is not None
        ctx.save_for_backward(hb, bias, labels, lse)
        return loss

    @staticmethod
    def backward(ctx, g):
        hb, bias, labels, lse = ctx.saved_tensors
        M, N, K = ctx.dims
        dev = g.device
        shadow = ctx.shadow
        ld = (N + 7) // 8 * 8
        dl = _empty_bf16(M, ld, dev=dev)
        if ld != N:
            dl[:, N:].zero_()
        ops.gemm(hb, shadow.get_nograd(), M, N, K, bias=bias, out_bf16=dl,
                 ce=dict(mode=2, labels=labels, lse=lse, row_grad=g.contiguous().float()))
        dbias = None
        if ctx.has_bias:  # the column-sum kernel works on whole 8-column groups: sum the padded width, drop the pad
            dbias = _zeros(ld, dev)
            ops.colsum_bf16(dl, M, ld, dbias)
            dbias = dbias[:N]
        dh = None
        if ctx.needs_input_grad[0]:
            if ctx.h_dtype == torch.bfloat16:
                dh = _empty_bf16(M, K, dev=dev)
                ops.gemm(dl, shadow.get_nograd(), M, K, N, b_mn=True, out_bf16=dh)
            else:
                dh = torch.empty(M, K, dtype=torch.float32, device=dev)
                ops.gemm(dl, shadow.get_nograd(), M, K, N, b_mn=True, out_f32=dh)
        wg = _wgrad(shadow, dl, hb, N, K, M)
        return (dh, dbias, None, None, *wg)


def vocab_cross_entropy(h, shadow, bias, labels):
    """Per-row cross entropy of the vocabulary projection of h [M, K] (labels int64 [M], negative = ignored -> 0)."""
    _note_uses(shadow)
    return _VocabCEFn.apply(h, bias, labels, shadow, *shadow.params)


# ----------------------------------------------------------------------------------------------
# LayerNorm (stand-alone, fp32 in -> fp32 out)
# ----------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        M, D = x2.shape
        y = torch.empty_like(x2)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        ops.layernorm_fwd(x2, w, b, eps, y_f32=y, mean=mean, rstd=rstd)
        ctx.save_for_backward(x2, w, mean, rstd)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        M, D = x2.shape
        dy2 = dy.reshape(M, D).contiguous()
        dx = torch.empty_like(x2)
        dw, db = _zeros(D, dy.device), _zeros(D, dy.device)
        ops.layernorm_bwd(dy2, x2, w, mean, rstd, dx, dw, db)
        return dx.view(dy.shape), dw, db, None


def layer_norm(x, w, b, eps):
    return _LayerNormFn.apply(x, w, b, eps)


# ----------------------------------------------------------------------------------------------
# BEiT block (models/beit2.py:191-209)
# ----------------------------------------------------------------------------------------------
class _BeitBlockFn(torch.autograd.Function):
    """x += dp·γ1·proj(attn(LN1(x)));  x += dp·γ2·fc2(GELU(fc1(LN2(x))))."""

    @staticmethod
    def forward(ctx, x, dp_scale, dp_scale2, blk, n1w, n1b, qkv_bias, table, projb, g1, n2w, n2b, fc1b, fc2b, g2, *weights):
        B, N, D = x.shape
        H = blk.attn.num_heads
        M = B * N
        dev = x.device
        x2 = x.contiguous().view(M, D)
        sh = blk._x2k
        mean1, rstd1 = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ln1 = _empty_bf16(M, D, dev=dev)
        ops.layernorm_fwd(x2, n1w, n1b, 1e-6, y_bf16=ln1, mean=mean1, rstd=rstd1)
        qkv = _empty_bf16(M, 3 * D, dev=dev)
        ops.gemm(ln1, sh["qkv"].get(), M, 3 * D, D, bias=qkv_bias, out_bf16=qkv)
        ldb = ops.pad32(N)
        bias_g = None
        if table is not None:
            bias_g = torch.empty(H, N, ldb, dtype=torch.float32, device=dev)
            ops.relpos_bias_gather(table, blk.attn.relative_position_index, N, H, bias_g)
        attn_o = _empty_bf16(M, D, dev=dev)
        lse = torch.empty(B, H, N, dtype=torch.float32, device=dev)
        scale = blk.attn.scale
        ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, scale, attn_o, lse, bias=bias_g)
        y1 = _empty_bf16(M, D, dev=dev)
        x1 = torch.empty(M, D, dtype=torch.float32, device=dev)
        ops.gemm(attn_o, sh["proj"].get(), M, D, D, bias=projb, preact_out=y1, gamma=g1, row_scale=dp_scale,
                 rows_per_scale=N, residual=x2, out_f32=x1)
        mean2, rstd2 = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ln2 = _empty_bf16(M, D, dev=dev)
        ops.layernorm_fwd(x1, n2w, n2b, 1e-6, y_bf16=ln2, mean=mean2, rstd=rstd2)
        Dh = sh["fc1"].total_rows
        hpre = _empty_bf16(M, Dh, dev=dev)
        act = _empty_bf16(M, Dh, dev=dev)
        # hpre receives GELU'(pre-activation) (evaluated next to GELU itself): the backward dgrad only multiplies by it
        ops.gemm(ln2, sh["fc1"].get(), M, Dh, D, bias=fc1b, preact_out=hpre, act=ACT_GELU_SAVE_GRAD, out_bf16=act)
        y2 = _empty_bf16(M, D, dev=dev)
        out = torch.empty(M, D, dtype=torch.float32, device=dev)
        ops.gemm(act, sh["fc2"].get(), M, D, Dh, bias=fc2b, preact_out=y2, gamma=g2, row_scale=dp_scale2,
                 rows_per_scale=N, residual=x1, out_f32=out)
        ctx.blk, ctx.dims, ctx.scale = blk, (B, N, D, H, Dh, ldb), scale
        ctx.bias_in_arena = qkv_bias is not None and not qkv_bias.requires_grad  # a view of the arena (GappedBias), not a torch.cat
        ctx.P = (n1w, n1b, table, projb, g1, n2w, n2b, fc1b, fc2b, g2)
        ctx.save_for_backward(x2, dp_scale, dp_scale2, n1w, table, g1, n2w, g2, mean1, rstd1, ln1, qkv, bias_g, attn_o, lse, y1, x1,
                              mean2, rstd2, ln2, hpre, act, y2)
        return out.view(B, N, D)

    @staticmethod
    def backward(ctx, dout):
        (x2, dp_scale, dp_scale2, n1w, table, g1, n2w, g2, mean1, rstd1, ln1, qkv, bias_g, attn_o, lse, y1, x1, mean2, rstd2, ln2,
         hpre, act, y2) = ctx.saved_tensors
        blk = ctx.blk
        sh = blk._x2k
        B, N, D, H, Dh, ldb = ctx.dims
        M = B * N
        dev = dout.device
        P_n1w, P_n1b, P_table, P_projb, P_g1, P_n2w, P_n2b, P_fc1b, P_fc2b, P_g2 = (_G(q, dev) for q in ctx.P)
        dx2 = dout.contiguous().view(M, D)
        # ---- MLP branch ----
        g2b = _empty_bf16(M, D, dev=dev)
        ops.scale_cast_colsum(dx2, M, D, g_bf16=g2b, gamma=g2, row_scale=dp_scale2, rows_per_scale=N,
                              y_bf16=y2 if g2 is not None else None, dbias=P_fc2b.buf, dgamma=P_g2.buf)
        dh = _empty_bf16(M, Dh, dev=dev)
        ops.gemm(g2b, sh["fc2"].get_nograd(), M, Dh, D, b_mn=True, act=ACT_MUL_AUX, aux=hpre, out_bf16=dh)
        wg_fc2 = _wgrad(sh["fc2"], g2b, act, D, Dh, M)
        del act, g2b
        ops.colsum_bf16(dh, M, Dh, P_fc1b.buf)
        dln2 = _empty_bf16(M, D, dev=dev)
        ops.gemm(dh, sh["fc1"].get_nograd(), M, D, Dh, b_mn=True, out_bf16=dln2)
        wg_fc1 = _wgrad(sh["fc1"], dh, ln2, Dh, D, M)
        del dh
        dx1 = torch.empty(M, D, dtype=torch.float32, device=dev)
        ops.layernorm_bwd(dln2, x1, n2w, mean2, rstd2, dx1, P_n2w.buf, P_n2b.buf, dx_residual=dx2)
        # ---- attention branch ----
        g1b = _empty_bf16(M, D, dev=dev)
        ops.scale_cast_colsum(dx1, M, D, g_bf16=g1b, gamma=g1, row_scale=dp_scale, rows_per_scale=N,
                              y_bf16=y1 if g1 is not None else None, dbias=P_projb.buf, dgamma=P_g1.buf)
        dattn = _empty_bf16(M, D, dev=dev)
        ops.gemm(g1b, sh["proj"].get_nograd(), M, D, D, b_mn=True, out_bf16=dattn)
        wg_proj = _wgrad(sh["proj"], g1b, attn_o, D, D, M)
        dqkv = _empty_bf16(M, 3 * D, dev=dev)
        ds_out = _empty_bf16(B, H, N, ldb, dev=dev) if table is not None else None
        ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, ctx.scale, attn_o, lse, dattn, dqkv[:, :D],
                     dqkv[:, D:2 * D], dqkv[:, 2 * D:], ds_out=ds_out, bias=bias_g)
        if table is not None:
            ops.relpos_bias_scatter(ds_out, B, H, N, blk.attn.relative_position_index, P_table.buf)
        bqv = sh.get("bqv")
        if bqv is not None and bqv.arena is not None and ctx.bias_in_arena:
            # q / v bias gradients: column sums straight into their slices of the flat gradient (K has no bias)
            gq, gv = bqv.grad_views()
            if bqv.params[0].requires_grad:
                ops.colsum_bf16(dqkv[:, :D], M, D, gq)
            if bqv.params[1].requires_grad:
                ops.colsum_bf16(dqkv[:, 2 * D:], M, D, gv)
            bqv.grads_done()
            d_qkvb = None
        else:
            d_qkvb = _zeros(3 * D, dev)
            ops.colsum_bf16(dqkv, M, 3 * D, d_qkvb)
        dln1 = _empty_bf16(M, D, dev=dev)
        ops.gemm(dqkv, sh["qkv"].get_nograd(), M, D, 3 * D, b_mn=True, out_bf16=dln1)
        wg_qkv = _wgrad(sh["qkv"], dqkv, ln1, 3 * D, D, M)
        dx = torch.empty(M, D, dtype=torch.float32, device=dev)
        ops.layernorm_bwd(dln1, x2, n1w, mean1, rstd1, dx, P_n1w.buf, P_n1b.buf, dx_residual=dx1)
        # weights were passed in the order qkv, proj, fc1, fc2
        nig = ctx.needs_input_grad
        return (dx.view(B, N, D), None, None, None, P_n1w.ret(), P_n1b.ret(), d_qkvb if (nig[6] and d_qkvb is not None) else None,
                P_table.ret(),
                P_projb.ret(), P_g1.ret(), P_n2w.ret(), P_n2b.ret(), P_fc1b.ret(), P_fc2b.ret(), P_g2.ret(), *wg_qkv, *wg_proj,
                *wg_fc1, *wg_fc2)


def beit_block(x, blk, dp_scale=None, dp_scale2=None):
    """x: [B, N, D] fp32.  dp_scale / dp_scale2: per-sample DropPath keep/(1-p) [B] fp32 (or None) of the attention and of
    the MLP branch — two independent draws, as the reference calls self.drop_path twice per block (beit2.py:204-207)."""
    a = blk.attn
    sh = blk._x2k
    # K has no bias (beit2.py:129): [q_bias | 0 | v_bias] is a view of the arena when there is one, else a torch.cat
    bqv = sh.get("bqv")
    if a.q_bias is None:
        qkv_bias = None
    elif bqv is not None and bqv.arena is not None:
        qkv_bias = bqv.get_f32()
        _note_uses(bqv)
    else:
        qkv_bias = torch.cat((a.q_bias, torch.zeros_like(a.v_bias), a.v_bias))
    weights = (*sh["qkv"].params, *sh["proj"].params, *sh["fc1"].params, *sh["fc2"].params)
    _note_uses(sh["qkv"], sh["proj"], sh["fc1"], sh["fc2"], blk.norm1.weight, blk.norm1.bias, a.relative_position_bias_table,
               a.proj.bias, blk.gamma_1, blk.norm2.weight, blk.norm2.bias, blk.mlp.fc1.bias, blk.mlp.fc2.bias, blk.gamma_2)
    return _BeitBlockFn.apply(x, dp_scale, dp_scale2, blk, blk.norm1.weight, blk.norm1.bias, qkv_bias, a.relative_position_bias_table,
                              a.proj.bias, blk.gamma_1, blk.norm2.weight, blk.norm2.bias, blk.mlp.fc1.bias,
                              blk.mlp.fc2.bias, blk.gamma_2, *weights)


# ----------------------------------------------------------------------------------------------
# BERT text / fusion layer (models/xbert.py:566-625), post-LN
# ----------------------------------------------------------------------------------------------
def _drop(p, train, n):
    if train and p > 0.0:
        seed, off = dropout_state.take(n)
        return p, seed, off
    return 0.0, 0, 0


def _pos_drop_path(mod, cfg, key, Bt, L, dev):
    """Per-row scale [Bt*L] of xbert's DropPath on one output module (models/xbert.py:518-535): floor(keep + U)/keep
    drawn once per sequence position and shared by the batch; None when inactive.  cfg['drop_path_scales'][key] may
    pin the [L] scale (tests)."""
    rate = getattr(getattr(mod, "drop_path", None), "drop_prob", None) or 0.0
    pinned = (cfg.get("drop_path_scales") or {}).get(key)
    if pinned is not None:
        scale = pinned.to(device=dev, dtype=torch.float32)
    elif cfg["train"] and rate > 0.0:
        keep = 1.0 - rate
        scale = torch.floor(keep + torch.rand(L, device=dev)) / keep
    else:
        return None
    return scale.repeat(Bt).contiguous()  # row m = b*L + l


def _dense_bwd(dy, s_in, lnw, mean, rstd, ds, P_lnw, P_lnb, g, dbias, drop, dp_rows, M, D):
    """Backward of LayerNorm(drop_path(dropout(dense(h))) + residual) up to the dense output: ds = LN'(dy) (the
    residual-stream gradient) and g = bf16(ds * dropout keep-scale * drop-path row scale), dbias += colsum(g)."""
    p, seed, off = drop
    if dp_rows is None:  # one fused pass
        ops.layernorm_bwd(dy, s_in, lnw, mean, rstd, ds, P_lnw.buf, P_lnb.buf, g_bf16=g, dbias=dbias, dropout_p=p,
                          dropout_seed=seed, dropout_offset=off)
    else:
        ops.layernorm_bwd(dy, s_in, lnw, mean, rstd, ds, P_lnw.buf, P_lnb.buf)
        ops.scale_cast_colsum(ds, M, D, g_bf16=g, row_scale=dp_rows, rows_per_scale=1, dropout_p=p, dropout_seed=seed,
                              dropout_offset=off, dbias=dbias)


class _BertLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, xb, enc, encb, layer, cfg, b_qkv, b_o, ln1w, ln1b, b_qc, b_kvc, b_oc, lncw, lncb, b_i, b_out, ln2w,
                ln2b, *weights):
        """x [Bt,L,D] fp32 (+ optional bf16 copy xb).  cfg: dict(self_mask, self_mask_3d, cross_mask, kv_index, n_kv,
        train, p_hidden, p_attn, eps).  enc [n_kv, Nk, Dv] fp32 / encb bf16 copy, or None (no cross-attention)."""
        Bt, L, D = x.shape
        M = Bt * L
        dev = x.device
        sh = layer._x2k
        H = layer.attention.self.num_attention_heads
        scale = 1.0 / math.sqrt(D // H)
        train, p_h, p_a, eps = cfg["train"], cfg["p_hidden"], cfg["p_attn"], cfg["eps"]
        x2 = x.contiguous().view(M, D)
        xb2 = xb.view(M, D) if xb is not None else ops.to_bf16(x2)
        sv = {}
        # ---- self-attention ----
        qkv = _empty_bf16(M, 3 * D, dev=dev)
        ops.gemm(xb2, sh["qkv"].get(), M, 3 * D, D, bias=b_qkv, out_bf16=qkv)
        ctx1 = _empty_bf16(M, D, dev=dev)
        lse1 = torch.empty(Bt, H, L, dtype=torch.float32, device=dev)
        d_a1 = _drop(p_a, train, Bt * H * L * ops.pad16(L))
        ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], Bt, H, L, L, scale, ctx1, lse1, mask=cfg["self_mask"],
                     mask_per_query=cfg["self_mask_3d"], dropout_p=d_a1[0], dropout_seed=d_a1[1], dropout_offset=d_a1[2])
        s1 = torch.empty(M, D, dtype=torch.float32, device=dev)
        d_h1 = _drop(p_h, train, M * D)
        dp1 = _pos_drop_path(layer.attention.output, cfg, "self", Bt, L, dev)
        ops.gemm(ctx1, sh["o"].get(), M, D, D, bias=b_o, dropout_p=d_h1[0], dropout_seed=d_h1[1], dropout_offset=d_h1[2],
                 row_scale=dp1, rows_per_scale=1, residual=x2, out_f32=s1)
        x1 = torch.empty(M, D, dtype=torch.float32, device=dev)
        x1b = _empty_bf16(M, D, dev=dev)
        m1, r1 = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.layernorm_fwd(s1, ln1w, ln1b, eps, y_bf16=x1b, y_f32=x1, mean=m1, rstd=r1)
        has_cross = enc is not None
        if has_cross:
            n_kv, Nk, Dv = enc.shape
            enc2 = encb.view(n_kv * Nk, Dv) if encb is not None else ops.to_bf16(enc.contiguous().view(n_kv * Nk, Dv))
            qc = _empty_bf16(M, D, dev=dev)
            ops.gemm(x1b, sh["qc"].get(), M, D, D, bias=b_qc, out_bf16=qc)
            kvc = _empty_bf16(n_kv * Nk, 2 * D, dev=dev)
            ops.gemm(enc2, sh["kvc"].get(), n_kv * Nk, 2 * D, Dv, bias=b_kvc, out_bf16=kvc)
            ctx2 = _empty_bf16(M, D, dev=dev)
            lse2 = torch.empty(Bt, H, L, dtype=torch.float32, device=dev)
            d_a2 = _drop(p_a, train, Bt * H * L * ops.pad16(Nk))
            ops.attn_fwd(qc, kvc[:, :D], kvc[:, D:], Bt, H, L, Nk, scale, ctx2, lse2, kv_index=cfg["kv_index"], n_kv=n_kv,
                         kv_groups=cfg.get("kv_groups"), mask=cfg["cross_mask"], mask_per_query=cfg.get("cross_mask_3d", False),
                         dropout_p=d_a2[0], dropout_seed=d_a2[1], dropout_offset=d_a2[2])
            if getattr(layer.crossattention.self, "save_attention", False):
                # Grad-CAM hooks of models/xbert.py:394-396: the (pre-dropout) map now, its gradient in backward
                layer.crossattention.self.save_attention_map(
                    ops.attn_probs(qc, kvc[:, :D], Bt, H, L, Nk, scale, lse2, kv_index=cfg["kv_index"], n_kv=n_kv,
                                   mask=cfg["cross_mask"], mask_per_query=cfg.get("cross_mask_3d", False)))
            s2 = torch.empty(M, D, dtype=torch.float32, device=dev)
            d_h2 = _drop(p_h, train, M * D)
            dp2 = _pos_drop_path(layer.crossattention.output, cfg, "cross", Bt, L, dev)
            ops.gemm(ctx2, sh["oc"].get(), M, D, D, bias=b_oc, dropout_p=d_h2[0], dropout_seed=d_h2[1],
                     dropout_offset=d_h2[2], row_scale=dp2, rows_per_scale=1, residual=x1, out_f32=s2)
            xa = torch.empty(M, D, dtype=torch.float32, device=dev)
            xab = _empty_bf16(M, D, dev=dev)
            mc, rc = torch.empty(M, device=dev), torch.empty(M, device=dev)
            ops.layernorm_fwd(s2, lncw, lncb, eps, y_bf16=xab, y_f32=xa, mean=mc, rstd=rc)
            sv.update(enc2=enc2, qc=qc, kvc=kvc, ctx2=ctx2, lse2=lse2, s2=s2, mc=mc, rc=rc, d_a2=d_a2, d_h2=d_h2,
                      dims_c=(n_kv, Nk, Dv), dp2=dp2)
        else:
            xa, xab = x1, x1b
        # ---- feed-forward ----
        Di = sh["i"].total_rows
        hpre = _empty_bf16(M, Di, dev=dev)
        act = _empty_bf16(M, Di, dev=dev)
        ops.gemm(xab, sh["i"].get(), M, Di, D, bias=b_i, preact_out=hpre, act=ACT_GELU_SAVE_GRAD, out_bf16=act)  # hpre = GELU''
        s3 = torch.empty(M, D, dtype=torch.float32, device=dev)
        d_h3 = _drop(p_h, train, M * D)
        dp3 = _pos_drop_path(layer.output, cfg, "ffn", Bt, L, dev)
        ops.gemm(act, sh["out"].get(), M, D, Di, bias=b_out, dropout_p=d_h3[0], dropout_seed=d_h3[1], dropout_offset=d_h3[2],
                 row_scale=dp3, rows_per_scale=1, residual=xa, out_f32=s3)
        y = torch.empty(M, D, dtype=torch.float32, device=dev)
        yb = _empty_bf16(M, D, dev=dev)
        m3, r3 = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.layernorm_fwd(s3, ln2w, ln2b, eps, y_bf16=yb, y_f32=y, mean=m3, rstd=r3)
        sv.update(xb2=xb2, qkv=qkv, ctx1=ctx1, lse1=lse1, s1=s1, m1=m1, r1=r1, x1b=x1b, xab=xab, hpre=hpre, act=act, s3=s3,
                  m3=m3, r3=r3, d_a1=d_a1, d_h1=d_h1, d_h3=d_h3, ln1w=ln1w, lncw=lncw, ln2w=ln2w, dp1=dp1, dp3=dp3)
        ctx.sv, ctx.layer, ctx.cfg, ctx.has_cross = sv, layer, cfg, has_cross
        ctx.P = (b_o, ln1w, ln1b, b_qc, b_oc, lncw, lncb, b_i, b_out, ln2w, ln2b)
        ctx.packed_bias = (b_qkv, b_kvc)
        ctx.dims = (Bt, L, D, H, Di, scale)
        ctx.enc_needs_grad = has_cross and enc.requires_grad
        if ctx.enc_needs_grad and cfg.get("fuse_d_enc"):
            cfg["_enc_total"] = cfg.get("_enc_total", 0) + 1
        ctx.mark_non_differentiable(yb)
        ctx.set_materialize_grads(False)  # no zero-filled [Bt, L, D] gradient for the non-differentiable bf16 copy
        return y.view(Bt, L, D), yb.view(Bt, L, D)

    @staticmethod
    def backward(ctx, dy, _dyb):
        sv, layer, cfg = ctx.sv, ctx.layer, ctx.cfg
        sh = layer._x2k
        Bt, L, D, H, Di, scale = ctx.dims
        M = Bt * L
        dev = dy.device
        dy2 = dy.contiguous().view(M, D)
        P_bo, P_ln1w, P_ln1b, P_bqc, P_boc, P_lncw, P_lncb, P_bi, P_bout, P_ln2w, P_ln2b = (
            _G(q if (ctx.has_cross or i not in (3, 4, 5, 6)) else None, dev) for i, q in enumerate(ctx.P))

        def packed_slot(key, tensor, n):
            """Packed bias (q|k|v or k|v): the arena's adjacent gradient views, else a zero buffer for autograd."""
            shd = sh.get(key)
            if tensor is not None and shd is not None and shd.arena is not None and not tensor.requires_grad:
                return shd.grad_sink().view(-1), shd
            return (_zeros(n, dev) if tensor is not None else None), None

        # ---- feed-forward ----
        # LayerNorm backward, dropout mask, bf16 cast and bias-gradient column sum in ONE pass over the rows
        ds3 = torch.empty(M, D, dtype=torch.float32, device=dev)
        g3 = _empty_bf16(M, D, dev=dev)
        _dense_bwd(dy2, sv["s3"], sv["ln2w"], sv["m3"], sv["r3"], ds3, P_ln2w, P_ln2b, g3, P_bout.buf, sv["d_h3"], sv["dp3"], M, D)
        dh = _empty_bf16(M, Di, dev=dev)
        ops.gemm(g3, sh["out"].get_nograd(), M, Di, D, b_mn=True, act=ACT_MUL_AUX, aux=sv["hpre"], out_bf16=dh)
        wg_out = _wgrad(sh["out"], g3, sv["act"], D, Di, M)
        ops.colsum_bf16(dh, M, Di, P_bi.buf)
        dxab = _empty_bf16(M, D, dev=dev)
        ops.gemm(dh, sh["i"].get_nograd(), M, D, Di, b_mn=True, out_bf16=dxab)
        wg_i = _wgrad(sh["i"], dh, sv["xab"], Di, D, M)
        del dh, g3
        d_bkvc = d_enc = None
        wg_qc, wg_kvc, wg_oc = [None] * len(sh["qc"].params) if "qc" in sh else [], \
            [None] * len(sh["kvc"].params) if "kvc" in sh else [], [None] * len(sh["oc"].params) if "oc" in sh else []
        if ctx.has_cross:
            n_kv, Nk, Dv = sv["dims_c"]
            ds2 = torch.empty(M, D, dtype=torch.float32, device=dev)
            g2 = _empty_bf16(M, D, dev=dev)
            _dense_bwd((ds3, dxab), sv["s2"], sv["lncw"], sv["mc"], sv["rc"], ds2, P_lncw, P_lncb, g2, P_boc.buf, sv["d_h2"],
                       sv["dp2"], M, D)
            dctx2 = _empty_bf16(M, D, dev=dev)
            ops.gemm(g2, sh["oc"].get_nograd(), M, D, D, b_mn=True, out_bf16=dctx2)
            wg_oc = _wgrad(sh["oc"], g2, sv["ctx2"], D, D, M)
            dqc = _empty_bf16(M, D, dev=dev)
            grouped = cfg.get("kv_groups") is not None  # grouped kernels return dK/dV per K/V source, already summed
            dkv_seq = _empty_bf16((n_kv if grouped else Bt) * Nk, 2 * D, dev=dev)
            p, seed, off = sv["d_a2"]
            kvc = sv["kvc"]
            ops.attn_bwd(sv["qc"], kvc[:, :D], kvc[:, D:], Bt, H, L, Nk, scale, sv["ctx2"], sv["lse2"], dctx2, dqc,
                         dkv_seq[:, :D], dkv_seq[:, D:], kv_index=cfg["kv_index"], n_kv=n_kv, kv_groups=cfg.get("kv_groups"),
                         mask=cfg["cross_mask"], mask_per_query=cfg.get("cross_mask_3d", False), dropout_p=p, dropout_seed=seed,
                         dropout_offset=off)
            if getattr(layer.crossattention.self, "save_attention", False):
                layer.crossattention.self.save_attn_gradients(
                    ops.attn_probs_grad(dctx2, kvc[:, D:], Bt, H, L, Nk, kv_index=cfg["kv_index"], n_kv=n_kv, dropout_p=p,
                                        dropout_seed=seed, dropout_offset=off))
            if cfg["kv_index"] is not None and not grouped:
                dkv = _empty_bf16(n_kv * Nk, 2 * D, dev=dev)
                ops.segment_sum_bf16(dkv_seq.view(Bt, Nk * 2 * D), cfg["kv_index"], n_kv, dkv.view(n_kv, Nk * 2 * D))
            else:
                dkv = dkv_seq
            d_bkvc, kv_shd = packed_slot("bkvc", ctx.packed_bias[1], 2 * D)
            ops.colsum_bf16(dkv, n_kv * Nk, 2 * D, d_bkvc)
            if kv_shd is not None:
                kv_shd.grads_done()
                d_bkvc = None
            if ctx.enc_needs_grad:
                if cfg.get("fuse_d_enc") and cfg.get("_enc_total", 0) > 1:
                    # every fusion layer of this encoder call reads the same image states: their gradients are summed in
                    # the dgrad GEMM's accumulate epilogue into ONE buffer, handed to autograd by the last layer to run
                    acc = cfg.get("_d_enc_acc")
                    if acc is None:
                        acc = cfg["_d_enc_acc"] = torch.empty(n_kv * Nk, Dv, dtype=torch.float32, device=dev)
                        ops.gemm(dkv, sh["kvc"].get_nograd(), n_kv * Nk, Dv, 2 * D, b_mn=True, out_f32=acc)
                    else:
                        ops.gemm(dkv, sh["kvc"].get_nograd(), n_kv * Nk, Dv, 2 * D, b_mn=True, out_f32=acc, accumulate=True)
                    cfg["_enc_left"] = cfg.get("_enc_left", cfg["_enc_total"]) - 1
                    if cfg["_enc_left"] == 0:
                        d_enc = acc.view(n_kv, Nk, Dv)
                        cfg["_d_enc_acc"], cfg["_enc_left"] = None, cfg["_enc_total"]
                else:
                    d_enc = torch.empty(n_kv * Nk, Dv, dtype=torch.float32, device=dev)
                    ops.gemm(dkv, sh["kvc"].get_nograd(), n_kv * Nk, Dv, 2 * D, b_mn=True, out_f32=d_enc)
                    d_enc = d_enc.view(n_kv, Nk, Dv)
            wg_kvc = _wgrad(sh["kvc"], dkv, sv["enc2"], 2 * D, Dv, n_kv * Nk)
            ops.colsum_bf16(dqc, M, D, P_bqc.buf)
            dx1b = _empty_bf16(M, D, dev=dev)
            ops.gemm(dqc, sh["qc"].get_nograd(), M, D, D, b_mn=True, out_bf16=dx1b)
            wg_qc = _wgrad(sh["qc"], dqc, sv["x1b"], D, D, M)
            res_f32, res_bf16 = ds2, dx1b
        else:
            res_f32, res_bf16 = ds3, dxab
        # ---- self-attention ----
        ds1 = torch.empty(M, D, dtype=torch.float32, device=dev)
        g1 = _empty_bf16(M, D, dev=dev)
        _dense_bwd((res_f32, res_bf16), sv["s1"], sv["ln1w"], sv["m1"], sv["r1"], ds1, P_ln1w, P_ln1b, g1, P_bo.buf, sv["d_h1"],
                   sv["dp1"], M, D)
        dctx1 = _empty_bf16(M, D, dev=dev)
        ops.gemm(g1, sh["o"].get_nograd(), M, D, D, b_mn=True, out_bf16=dctx1)
        wg_o = _wgrad(sh["o"], g1, sv["ctx1"], D, D, M)
        dqkv = _empty_bf16(M, 3 * D, dev=dev)
        qkv = sv["qkv"]
        p, seed, off = sv["d_a1"]
        ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], Bt, H, L, L, scale, sv["ctx1"], sv["lse1"], dctx1,
                     dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], mask=cfg["self_mask"],
                     mask_per_query=cfg["self_mask_3d"], dropout_p=p, dropout_seed=seed, dropout_offset=off)
        d_bqkv, qkv_shd = packed_slot("bqkv", ctx.packed_bias[0], 3 * D)
        ops.colsum_bf16(dqkv, M, 3 * D, d_bqkv)
        if qkv_shd is not None:
            qkv_shd.grads_done()
            d_bqkv = None
        dx = torch.empty(M, D, dtype=torch.float32, device=dev)
        ops.gemm(dqkv, sh["qkv"].get_nograd(), M, D, 3 * D, b_mn=True, residual=ds1, out_f32=dx)
        wg_qkv = _wgrad(sh["qkv"], dqkv, sv["xb2"], 3 * D, D, M)
        ctx.sv = None
        # weight order: qkv(3), o, [qc, kvc(2), oc], i, out
        return (dx.view(Bt, L, D), None, d_enc, None, None, None, d_bqkv, P_bo.ret(), P_ln1w.ret(), P_ln1b.ret(), P_bqc.ret(),
                d_bkvc, P_boc.ret(), P_lncw.ret(), P_lncb.ret(), P_bi.ret(), P_bout.ret(), P_ln2w.ret(), P_ln2b.ret(),
                *wg_qkv, *wg_o, *wg_qc, *wg_kvc, *wg_oc, *wg_i, *wg_out)


def bert_layer(x, xb, layer, cfg, enc=None, encb=None):
    """One BertLayer.  Returns (y fp32 [Bt,L,D], y bf16 copy)."""
    sh = layer._x2k
    at, so = layer.attention.self, layer.attention.output

    def packed(key, *biases):
        """fp32 packed bias: a plain view of the arena (members are adjacent) or a differentiable torch.cat."""
        shd = sh[key]
        if shd.arena is not None:
            return shd.get_f32()
        return torch.cat(biases)

    b_qkv = packed("bqkv", at.query.bias, at.key.bias, at.value.bias)
    use_cross = enc is not None and layer.has_cross_attention
    uses = [sh["qkv"], sh["o"], sh["i"], sh["out"], sh["bqkv"], so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias,
            layer.intermediate.dense.bias, layer.output.dense.bias, layer.output.LayerNorm.weight, layer.output.LayerNorm.bias]
    if use_cross:
        ca, co = layer.crossattention.self, layer.crossattention.output
        b_qc, b_kvc, b_oc = ca.query.bias, packed("bkvc", ca.key.bias, ca.value.bias), co.dense.bias
        lncw, lncb = co.LayerNorm.weight, co.LayerNorm.bias
        uses += [sh["qc"], sh["kvc"], sh["oc"], sh["bkvc"], b_qc, b_oc, lncw, lncb]
    else:
        b_qc = b_kvc = b_oc = lncw = lncb = None
        enc = encb = None
    weights = [*sh["qkv"].params, *sh["o"].params]
    if "qc" in sh:
        weights += [*sh["qc"].params, *sh["kvc"].params, *sh["oc"].params]
    weights += [*sh["i"].params, *sh["out"].params]
    _note_uses(*uses)
    return _BertLayerFn.apply(x, xb, enc, encb, layer, cfg, b_qkv, so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias,
                              b_qc, b_kvc, b_oc, lncw, lncb, layer.intermediate.dense.bias, layer.output.dense.bias,
                              layer.output.LayerNorm.weight, layer.output.LayerNorm.bias, *weights)


# ----------------------------------------------------------------------------------------------
# BERT layer, generation paths (inference only): history_states / past_key_value / use_cache
# ----------------------------------------------------------------------------------------------
def _kv_rows(t, B, H):
    """Cached keys/values in the reference layout [B, H, L, 64] -> token-major bf16 rows [B, L, H*64]."""
    L = t.shape[2]
    return t.permute(0, 2, 1, 3).reshape(B, L, H * t.shape[3]).to(torch.bfloat16)


def _kv_heads(rows, B, H):
    """token-major [B, L, H*64] -> reference cache layout [B, H, L, 64] (a view)."""
    L = rows.shape[1]
    return rows.view(B, L, H, -1).permute(0, 2, 1, 3)


@torch.no_grad()
def bert_layer_decode(x, layer, cfg, enc=None, encb=None, history=None, past_kv=None, return_probs=False):
    """One BertLayer forward for the generation paths of models/xbert.py:322-415 (no autograd graph):

    * `history` [B, Lh, D] — the layer's cached INPUT states (model_generation.py:180-188): keys/values are projected
      from cat(history, x) while queries come from x only (xbert.py:349-353);
    * `past_kv` (K, V) each [B, H, Lp, 64] — HF-style cache: new keys/values are appended (xbert.py:355-359).

    cfg['self_mask'] covers all Lk = Lh|Lp + Lq keys.  Returns (y fp32 [B,Lq,D], (K, V) in the cache layout); with
    return_probs also (self-attention map, cross-attention map or None), fp32 [B,H,Lq,Lk], pre-dropout — the tensors the
    reference returns under output_attentions (xbert.py:392,410)."""
    if history is not None and past_kv is not None:
        raise ValueError("history_states and past_key_value are mutually exclusive (xbert.py:350)")
    B, Lq, D = x.shape
    M = B * Lq
    dev = x.device
    sh = layer._x2k
    at, so = layer.attention.self, layer.attention.output
    H = at.num_attention_heads
    scale = 1.0 / math.sqrt(D // H)
    eps = cfg["eps"]
    w_qkv = sh["qkv"].get()
    b_qkv = sh["bqkv"].get_f32() if sh["bqkv"].arena is not None else torch.cat((at.query.bias, at.key.bias, at.value.bias))
    x2 = x.contiguous().view(M, D).float()
    xb2 = ops.to_bf16(x2)
    if history is not None:
        Lh = history.shape[1]
        xs = ops.to_bf16(torch.cat((history.float(), x.float()), dim=1).contiguous().view(B * (Lh + Lq), D))
        q = _empty_bf16(M, D, dev=dev)
        ops.gemm(xb2, w_qkv[:D], M, D, D, bias=b_qkv[:D], out_bf16=q)
        kv = _empty_bf16(B * (Lh + Lq), 2 * D, dev=dev)
        ops.gemm(xs, w_qkv[D:], B * (Lh + Lq), 2 * D, D, bias=b_qkv[D:], out_bf16=kv)
        Lk = Lh + Lq
        k_rows, v_rows = kv[:, :D], kv[:, D:]
    else:
        qkv = _empty_bf16(M, 3 * D, dev=dev)
        ops.gemm(xb2, w_qkv, M, 3 * D, D, bias=b_qkv, out_bf16=qkv)
        q = qkv[:, :D]
        if past_kv is not None:
            pk, pv = _kv_rows(past_kv[0], B, H), _kv_rows(past_kv[1], B, H)
            Lk = pk.shape[1] + Lq
            kv = torch.cat((torch.cat((pk, qkv.view(B, Lq, 3 * D)[:, :, D:2 * D]), dim=1),
                            torch.cat((pv, qkv.view(B, Lq, 3 * D)[:, :, 2 * D:]), dim=1)), dim=2).view(B * Lk, 2 * D)
            k_rows, v_rows = kv[:, :D], kv[:, D:]
        else:
            Lk = Lq
            kv = qkv[:, D:]
            k_rows, v_rows = qkv[:, D:2 * D], qkv[:, 2 * D:]
    present = (_kv_heads(k_rows.reshape(B, Lk, D), B, H), _kv_heads(v_rows.reshape(B, Lk, D), B, H))
    ctx1 = _empty_bf16(M, D, dev=dev)
    lse1 = torch.empty(B, H, Lq, dtype=torch.float32, device=dev)
    ops.attn_fwd(q, k_rows, v_rows, B, H, Lq, Lk, scale, ctx1, lse1, n_kv=B, mask=cfg["self_mask"],
                 mask_per_query=cfg["self_mask_3d"])
    p_self = p_cross = None
    if return_probs:
        p_self = ops.attn_probs(q, k_rows, B, H, Lq, Lk, scale, lse1, n_kv=B, mask=cfg["self_mask"],
                                mask_per_query=cfg["self_mask_3d"])
    s1 = torch.empty(M, D, dtype=torch.float32, device=dev)
    ops.gemm(ctx1, sh["o"].get(), M, D, D, bias=so.dense.bias, residual=x2, out_f32=s1)
    x1 = torch.empty(M, D, dtype=torch.float32, device=dev)
    x1b = _empty_bf16(M, D, dev=dev)
    ops.layernorm_fwd(s1, so.LayerNorm.weight, so.LayerNorm.bias, eps, y_bf16=x1b, y_f32=x1)
    if enc is not None and layer.has_cross_attention:
        ca, co = layer.crossattention.self, layer.crossattention.output
        n_kv, Nk, Dv = enc.shape
        enc2 = encb.view(n_kv * Nk, Dv) if encb is not None else ops.to_bf16(enc.contiguous().view(n_kv * Nk, Dv))
        b_kvc = sh["bkvc"].get_f32() if sh["bkvc"].arena is not None else torch.cat((ca.key.bias, ca.value.bias))
        qc = _empty_bf16(M, D, dev=dev)
        ops.gemm(x1b, sh["qc"].get(), M, D, D, bias=ca.query.bias, out_bf16=qc)
        kvc = _empty_bf16(n_kv * Nk, 2 * D, dev=dev)
        ops.gemm(enc2, sh["kvc"].get(), n_kv * Nk, 2 * D, Dv, bias=b_kvc, out_bf16=kvc)
        ctx2 = _empty_bf16(M, D, dev=dev)
        lse2 = torch.empty(B, H, Lq, dtype=torch.float32, device=dev)
        ops.attn_fwd(qc, kvc[:, :D], kvc[:, D:], B, H, Lq, Nk, scale, ctx2, lse2, kv_index=cfg["kv_index"], n_kv=n_kv,
                     kv_groups=cfg.get("kv_groups"), mask=cfg["cross_mask"], mask_per_query=cfg.get("cross_mask_3d", False))
        if return_probs:
            p_cross = ops.attn_probs(qc, kvc[:, :D], B, H, Lq, Nk, scale, lse2, kv_index=cfg["kv_index"], n_kv=n_kv,
                                     mask=cfg["cross_mask"], mask_per_query=cfg.get("cross_mask_3d", False))
        s2 = torch.empty(M, D, dtype=torch.float32, device=dev)
        ops.gemm(ctx2, sh["oc"].get(), M, D, D, bias=co.dense.bias, residual=x1, out_f32=s2)
        xa = torch.empty(M, D, dtype=torch.float32, device=dev)
        xab = _empty_bf16(M, D, dev=dev)
        ops.layernorm_fwd(s2, co.LayerNorm.weight, co.LayerNorm.bias, eps, y_bf16=xab, y_f32=xa)
    else:
        xa, xab = x1, x1b
    Di = sh["i"].total_rows
    act = _empty_bf16(M, Di, dev=dev)
    ops.gemm(xab, sh["i"].get(), M, Di, D, bias=layer.intermediate.dense.bias, act=ACT_GELU, out_bf16=act)
    s3 = torch.empty(M, D, dtype=torch.float32, device=dev)
    ops.gemm(act, sh["out"].get(), M, D, Di, bias=layer.output.dense.bias, residual=xa, out_f32=s3)
    y = torch.empty(M, D, dtype=torch.float32, device=dev)
    ops.layernorm_fwd(s3, layer.output.LayerNorm.weight, layer.output.LayerNorm.bias, eps, y_f32=y)
    if return_probs:
        return y.view(B, Lq, D), present, (p_self, p_cross)
    return y.view(B, Lq, D), present


@torch.no_grad()
def beit_attention_map(x, blk):
    """Pre-dropout attention probabilities [B,H,N,N] of one BEiT block for input x (the `attn_prob` the reference's
    Attention.forward returns, models/beit2.py:152,166): LN1 -> QKV -> lse by the fused forward -> streaming map kernel."""
    B, N, D = x.shape
    M, dev = B * N, x.device
    a = blk.attn
    H = a.num_heads
    ln1 = _empty_bf16(M, D, dev=dev)
    ops.layernorm_fwd(x.contiguous().view(M, D).float(), blk.norm1.weight, blk.norm1.bias, 1e-6, y_bf16=ln1)
    qkv_bias = torch.cat((a.q_bias, torch.zeros_like(a.v_bias), a.v_bias)) if a.q_bias is not None else None
    qkv = _empty_bf16(M, 3 * D, dev=dev)
    ops.gemm(ln1, blk._x2k["qkv"].get_nograd(), M, 3 * D, D, bias=qkv_bias, out_bf16=qkv)
    bias_g = None
    if a.relative_position_bias_table is not None:
        bias_g = torch.empty(H, N, ops.pad32(N), dtype=torch.float32, device=dev)
        ops.relpos_bias_gather(a.relative_position_bias_table, a.relative_position_index, N, H, bias_g)
    o = _empty_bf16(M, D, dev=dev)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=dev)
    ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, N, N, a.scale, o, lse, bias=bias_g)
    return ops.attn_probs(qkv[:, :D], qkv[:, D:2 * D], B, H, N, N, a.scale, lse, bias=bias_g)
