"""TEST INFRASTRUCTURE — CPU restatement (plain torch fp32, functional, no nn.Module) of the
reference algorithm on the hot path.  Each function cites the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module — as the checker / the reported CPU baseline, never as the product path
(x2vlm_b200/ must not import oracle/).

Pinning: the reference ships no tests or golden vectors.  This restatement is
pinned against the reference ITSELF, run unmodified through oracle/ref_shim.py in the build
container: tests/test_oracle_vs_reference.py compares every function below with the reference's
modules on identical state_dicts and seeded inputs, and oracle/make_golden.py stores reference
outputs under tests/golden/ so the comparison also runs where /root/reference is absent.

All functions take `sd`, a dict {reference state_dict key: tensor}, so weights are interchangeable
with the reference and with x2vlm_b200's drop-in modules.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# BEiT-2 vision encoder (models/beit2.py)
# ----------------------------------------------------------------------------------------------
def relative_position_index(window):
    """int64 [N,N] index into the rel-pos table, N = Wh*Ww + 1 (models/beit2.py:97-111)."""
    wh, ww = window
    ch, cw = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    coords = torch.stack([ch.reshape(-1), cw.reshape(-1)])  # 2, P
    rel = coords[:, :, None] - coords[:, None, :]  # 2, P, P
    idx_pp = (rel[0] + wh - 1) * (2 * ww - 1) + (rel[1] + ww - 1)
    num_rel = (2 * wh - 1) * (2 * ww - 1) + 3
    n = wh * ww + 1
    idx = torch.zeros(n, n, dtype=torch.int64)
    idx[1:, 1:] = idx_pp
    idx[0, :] = num_rel - 3
    idx[:, 0] = num_rel - 2
    idx[0, 0] = num_rel - 1
    return idx


def beit_rel_pos_bias(sd, pfx):
    """[H,N,N] bias gathered from the per-block table (models/beit2.py:138-143)."""
    table = sd[pfx + "attn.relative_position_bias_table"]
    index = sd[pfx + "attn.relative_position_index"]
    n = index.shape[0]
    return table[index.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1).contiguous()


def beit_attention(x, sd, pfx, num_heads):
    """models/beit2.py:125-166 (eval: attn_drop = proj_drop = 0). Returns (out, attn_prob)."""
    B, N, C = x.shape
    d = C // num_heads
    qb, vb = sd[pfx + "attn.q_bias"], sd[pfx + "attn.v_bias"]
    bias = torch.cat([qb, torch.zeros_like(vb), vb])  # K has no bias (:129)
    qkv = F.linear(x, sd[pfx + "attn.qkv.weight"], bias).reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]  # q scaled before q@k^T (:135-136)
    attn = q @ k.transpose(-2, -1) + beit_rel_pos_bias(sd, pfx).unsqueeze(0)
    prob = attn.softmax(dim=-1)
    out = (prob @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, sd[pfx + "attn.proj.weight"], sd[pfx + "attn.proj.bias"]), prob


def beit_mlp(x, sd, pfx):
    """fc2(GELU_erf(fc1(x))) — models/beit2.py:61-68."""
    h = F.gelu(F.linear(x, sd[pfx + "mlp.fc1.weight"], sd[pfx + "mlp.fc1.bias"]))
    return F.linear(h, sd[pfx + "mlp.fc2.weight"], sd[pfx + "mlp.fc2.bias"])


def beit_block(x, sd, pfx, num_heads, drop_path_scale=None, drop_path_scale2=None):
    """Pre-LN block with LayerScale; drop_path_scale / drop_path_scale2 = per-sample keep/(1-p) [B] (or None) of the
    attention / MLP branch — the reference draws them independently (models/beit2.py:204-207; timm drop_path is per
    sample).  Passing only the first applies it to both branches."""
    dp = 1.0 if drop_path_scale is None else drop_path_scale.view(-1, 1, 1)
    dp2 = dp if drop_path_scale2 is None else drop_path_scale2.view(-1, 1, 1)
    C = x.shape[-1]
    y, prob = beit_attention(F.layer_norm(x, (C,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], 1e-6), sd, pfx,
                             num_heads)
    x = x + dp * (sd[pfx + "gamma_1"] * y)
    y = beit_mlp(F.layer_norm(x, (C,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], 1e-6), sd, pfx)
    x = x + dp2 * (sd[pfx + "gamma_2"] * y)
    return x, prob


def beit_patch_embed(image, sd, pfx):
    """Conv2d(k = s = patch) -> flatten -> transpose (models/beit2.py:227-232)."""
    w = sd[pfx + "patch_embed.proj.weight"]
    x = F.conv2d(image, w, sd[pfx + "patch_embed.proj.bias"], stride=w.shape[-1])
    return x.flatten(2).transpose(1, 2)


def vision_forward(image, sd, pfx, depth, num_heads, idx_to_group_img=None, image_atts=None, drop_path_scales=None):
    """VisionTransformer.forward with use_abs_pos_emb=False, use_mean_pooling=True
    (models/beit2.py:378-436).  Returns embeds [B,N,D], or (region_embeds, full_embeds)."""
    x = beit_patch_embed(image, sd, pfx)
    B = x.shape[0]
    x = torch.cat([sd[pfx + "cls_token"].expand(B, -1, -1), x], dim=1)
    for i in range(depth):
        x, _ = beit_block(x, sd, "%sblocks.%d." % (pfx, i), num_heads,
                          None if drop_path_scales is None else drop_path_scales[i])
    x = x[:, 1:]  # cls output is dropped (:409)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), sd[pfx + "fc_norm.weight"], sd[pfx + "fc_norm.bias"], 1e-6)
    x_cls = x.mean(dim=1, keepdim=True)
    full = torch.cat([x_cls, x], dim=1)
    if idx_to_group_img is None:
        return full
    x_bs = x[idx_to_group_img]  # gather per region sample (:430)
    w = image_atts[:, 1:].unsqueeze(2).to(x.dtype)
    x_bs_cls = (w * x_bs).sum(dim=1, keepdim=True) / w.sum(dim=1, keepdim=True)
    return torch.cat([x_bs_cls, x_bs], dim=1), full


# ----------------------------------------------------------------------------------------------
# BERT text / fusion encoder (models/xbert.py)
# ----------------------------------------------------------------------------------------------
def bert_embeddings(input_ids, sd, pfx):
    """word + token_type(0) + position[0:L] -> LayerNorm(1e-12) (models/xbert.py:189-216; eval)."""
    L = input_ids.shape[1]
    e = sd[pfx + "word_embeddings.weight"][input_ids] + sd[pfx + "token_type_embeddings.weight"][0] \
        + sd[pfx + "position_embeddings.weight"][:L]
    return F.layer_norm(e, (e.shape[-1],), sd[pfx + "LayerNorm.weight"], sd[pfx + "LayerNorm.bias"], 1e-12)


def extended_self_mask(attention_mask, dtype=torch.float32):
    """(1 - m) * -10000 broadcast to [B,1,1|L,L] (models/xbert.py:1013-1073, encoder case)."""
    m = attention_mask[:, None, :, :] if attention_mask.dim() == 3 else attention_mask[:, None, None, :]
    return (1.0 - m.to(dtype)) * -10000.0


def extended_cross_mask(encoder_attention_mask, dtype=torch.float32):
    """HF invert_attention_mask: masked keys get a large negative additive term
    (models/xbert.py:1165-1170; identical results to the pinned 4.12.5 whenever a key is unmasked)."""
    m = encoder_attention_mask[:, None, :, :] if encoder_attention_mask.dim() == 3 else encoder_attention_mask[:, None, None, :]
    return (1.0 - m.to(dtype)) * torch.finfo(dtype).min


def bert_attention_core(hidden, kv_src, ext_mask, sd, pfx, num_heads, train=False, p_attn=0.1):
    """query/key/value Linears, scores/sqrt(d) + mask, softmax, (dropout), PV, merge heads
    (models/xbert.py:322-415).  Returns (context, probs)."""
    B, L, C = hidden.shape
    d = C // num_heads

    def split(t):
        return t.reshape(t.shape[0], t.shape[1], num_heads, d).permute(0, 2, 1, 3)

    q = split(F.linear(hidden, sd[pfx + "query.weight"], sd[pfx + "query.bias"]))
    k = split(F.linear(kv_src, sd[pfx + "key.weight"], sd[pfx + "key.bias"]))
    v = split(F.linear(kv_src, sd[pfx + "value.weight"], sd[pfx + "value.bias"]))
    scores = q @ k.transpose(-1, -2) / math.sqrt(d)
    if ext_mask is not None:
        scores = scores + ext_mask
    probs = scores.softmax(dim=-1)
    pd = F.dropout(probs, p_attn, training=train)
    ctx = (pd @ v).permute(0, 2, 1, 3).reshape(B, L, C)
    return ctx, probs


def drop_path_scale(L, drop_prob):
    """xbert's DropPath mask (models/xbert.py:518-535) as a per-position scale [L]: floor(keep + U(0,1)) / keep, ONE
    draw per sequence position shared by the whole batch (shape (1, L, 1) in the reference).  Consumes torch's global
    RNG exactly like the reference's torch.rand((1, L, 1))."""
    keep = 1.0 - drop_prob
    return torch.floor(keep + torch.rand((1, L, 1))).view(L) / keep


def bert_output_ln(h, residual, sd, pfx, train=False, p_hidden=0.1, pos_scale=None):
    """LayerNorm(drop_path(dropout(dense(h))) + residual), eps 1e-12 (models/xbert.py:427-431, :511-515);
    pos_scale [L] = the DropPath scale per sequence position (None: DropPath off)."""
    y = F.dropout(F.linear(h, sd[pfx + "dense.weight"], sd[pfx + "dense.bias"]), p_hidden, training=train)
    if pos_scale is not None:
        y = y * pos_scale[None, :, None]
    return F.layer_norm(y + residual, (y.shape[-1],), sd[pfx + "LayerNorm.weight"], sd[pfx + "LayerNorm.bias"], 1e-12)


def bert_layer(hidden, self_mask, sd, pfx, num_heads, enc_hidden=None, cross_mask=None, train=False, p_attn=0.1, p_hidden=0.1,
               dp_scales=None):
    """self-attn -> (cross-attn if the layer has one AND encoder states are given) -> FFN
    (models/xbert.py:566-625).  dp_scales: {'self','cross','ffn'} -> per-position DropPath scale [L] of the three
    output modules (None: off)."""
    dp = dp_scales or {}
    ctx, _ = bert_attention_core(hidden, hidden, self_mask, sd, pfx + "attention.self.", num_heads, train, p_attn)
    x = bert_output_ln(ctx, hidden, sd, pfx + "attention.output.", train, p_hidden, dp.get("self"))
    if enc_hidden is not None and (pfx + "crossattention.self.query.weight") in sd:
        ctx, _ = bert_attention_core(x, enc_hidden, cross_mask, sd, pfx + "crossattention.self.", num_heads, train, p_attn)
        x = bert_output_ln(ctx, x, sd, pfx + "crossattention.output.", train, p_hidden, dp.get("cross"))
    inter = F.gelu(F.linear(x, sd[pfx + "intermediate.dense.weight"], sd[pfx + "intermediate.dense.bias"]))
    return bert_output_ln(inter, x, sd, pfx + "output.", train, p_hidden, dp.get("ffn"))


def bert_encoder(hidden, attention_mask, sd, pfx, num_heads, fusion_layer, num_layers, mode="multi_modal",
                 enc_hidden=None, enc_mask=None, train=False):
    """Layer range by mode (models/xbert.py:674-686) over BertLayer."""
    start, end = {"text": (0, fusion_layer), "fusion": (fusion_layer, num_layers),
                  "multi_modal": (0, num_layers)}[mode]
    self_mask = extended_self_mask(attention_mask, hidden.dtype)
    cross_mask = None
    if enc_hidden is not None:
        if enc_mask is None:
            enc_mask = torch.ones(enc_hidden.shape[:2])
        cross_mask = extended_cross_mask(enc_mask, hidden.dtype)
    for i in range(start, end):
        hidden = bert_layer(hidden, self_mask, sd, "%sencoder.layer.%d." % (pfx, i), num_heads, enc_hidden, cross_mask,
                            train)
    return hidden


def bert_model(sd, pfx, num_heads, fusion_layer, num_layers, input_ids=None, encoder_embeds=None, attention_mask=None,
               enc_hidden=None, enc_mask=None, mode="multi_modal", train=False):
    """BertModel.forward (models/xbert.py:1075-1220), encoder (non-decoder) case."""
    if encoder_embeds is None:
        hidden = bert_embeddings(input_ids, sd, pfx + "embeddings.")
        if train:
            hidden = F.dropout(hidden, 0.1, training=True)
    else:
        hidden = encoder_embeds
    return bert_encoder(hidden, attention_mask, sd, pfx, num_heads, fusion_layer, num_layers, mode, enc_hidden, enc_mask,
                        train)


def mlm_head(seq, sd, pfx):
    """transform (dense + GELU + LN) then tied decoder + bias (models/xbert.py:785-834)."""
    h = F.gelu(F.linear(seq, sd[pfx + "transform.dense.weight"], sd[pfx + "transform.dense.bias"]))
    h = F.layer_norm(h, (h.shape[-1],), sd[pfx + "transform.LayerNorm.weight"], sd[pfx + "transform.LayerNorm.bias"], 1e-12)
    return F.linear(h, sd[pfx + "decoder.weight"], sd[pfx + "bias"])


def gather_by_pos(seq, pos):
    """models/xbert.py:1647 gather_seq_out_by_pos."""
    return torch.gather(seq, 1, pos.unsqueeze(2).expand(-1, -1, seq.size(-1)))


# ----------------------------------------------------------------------------------------------
# X-VLM losses (models/xvlm.py, models/model_pretrain.py, models/box_ops.py)
# ----------------------------------------------------------------------------------------------
def mlp_head(x, sd, pfx):
    """Linear - LayerNorm - GELU - Linear (models/xvlm.py:163-169)."""
    h = F.linear(x, sd[pfx + "0.weight"], sd[pfx + "0.bias"])
    h = F.gelu(F.layer_norm(h, (h.shape[-1],), sd[pfx + "1.weight"], sd[pfx + "1.bias"], 1e-5))
    return F.linear(h, sd[pfx + "3.weight"], sd[pfx + "3.bias"])


def get_features(image_embeds, text_embeds, sd):
    """models/xvlm.py:785-792."""
    fi = F.normalize(F.linear(image_embeds[:, 0], sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1)
    ft = F.normalize(F.linear(text_embeds[:, 0], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1)
    return fi, ft


def contrastive_loss(image_feat, text_feat, temp):
    """Single-rank ITC, idx=None (models/xvlm.py:794-826)."""
    logits = image_feat @ text_feat.t() / temp
    labels = torch.arange(logits.shape[0])
    return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2


def hard_negative_weights(image_feat, text_feat, temp):
    """softmax(sim) + 1e-5 with the diagonal zeroed (models/xvlm.py:828-845)."""
    with torch.no_grad():
        w_i2t = F.softmax(image_feat @ text_feat.t() / temp, dim=1) + 1e-5
        w_t2i = F.softmax(text_feat @ image_feat.t() / temp, dim=1) + 1e-5
        w_i2t.fill_diagonal_(0)
        w_t2i.fill_diagonal_(0)
    return w_i2t, w_t2i


def box_cxcywh_to_xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def giou_diag(b1, b2):
    """diag of generalized_box_iou (models/box_ops.py:27-58), xyxy boxes."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt, rb = torch.max(b1[:, :2], b2[:, :2]), torch.min(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = a1 + a2 - inter
    iou = inter / union
    lt, rb = torch.min(b1[:, :2], b2[:, :2]), torch.max(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[:, 0] * wh[:, 1]
    return iou - (area - union) / area


def bbox_loss(output_coord, target_bbox, is_image=None):
    """L1 + GIoU with the degenerate-box guard (models/xvlm.py:927-957)."""
    l1 = F.l1_loss(output_coord, target_bbox, reduction="none")
    b1, b2 = box_cxcywh_to_xyxy(output_coord), box_cxcywh_to_xyxy(target_bbox)
    if (b1[:, 2:] < b1[:, :2]).any() or (b2[:, 2:] < b2[:, :2]).any():
        giou = torch.zeros(output_coord.size(0))
    else:
        giou = 1 - giou_diag(b1, b2)
    if is_image is None:
        n = target_bbox.size(0)
    else:
        n = torch.sum(1 - is_image)
        l1 = l1 * (1 - is_image.view(-1, 1))
        giou = giou * (1 - is_image)
    return l1.sum() / n, giou.sum() / n



# ------------------------------------------------------------------------------------------------
# generation paths: causal decoder, HF-style key/value cache, captioning history_states
# ------------------------------------------------------------------------------------------------
def decoder_self_mask(attention_mask, Lq):
    """get_extended_attention_mask with is_decoder=True (models/xbert.py:1013-1073): a 2-D key mask [B, Lk] is
    combined with a causal mask over the last Lq positions (cached prefix positions are all visible); a 3-D mask
    [B, Lq, Lk] is taken as is.  Returns the additive mask [B, 1, Lq, Lk]."""
    if attention_mask.dim() == 3:
        ext = attention_mask[:, None, :, :]
    else:
        B, Lk = attention_mask.shape
        ids = torch.arange(Lq)
        causal = (ids[None, None, :].repeat(B, Lq, 1) <= ids[None, :, None]).to(attention_mask.dtype)
        if Lq < Lk:
            causal = torch.cat([torch.ones(B, Lq, Lk - Lq, dtype=causal.dtype), causal], dim=-1)
        ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
    return (1.0 - ext.float()) * -10000.0


def bert_layer_cached(hidden, self_mask, sd, pfx, num_heads, enc_hidden=None, cross_mask=None, history=None, past=None):
    """BertLayer with the key/value sources of the generation paths (models/xbert.py:349-359): keys/values from
    cat(history, hidden) (history_states) or cat(past K/V, new K/V) (past_key_value).  Returns (out, (K, V))."""
    B, L, C = hidden.shape
    d = C // num_heads
    a = pfx + "attention.self."

    def split(t):
        return t.reshape(t.shape[0], t.shape[1], num_heads, d).permute(0, 2, 1, 3)

    q = split(F.linear(hidden, sd[a + "query.weight"], sd[a + "query.bias"]))
    src = torch.cat((history, hidden), dim=1) if history is not None else hidden
    k = split(F.linear(src, sd[a + "key.weight"], sd[a + "key.bias"]))
    v = split(F.linear(src, sd[a + "value.weight"], sd[a + "value.bias"]))
    if past is not None:
        assert history is None
        k, v = torch.cat([past[0], k], dim=2), torch.cat([past[1], v], dim=2)
    probs = (q @ k.transpose(-1, -2) / math.sqrt(d) + self_mask).softmax(dim=-1)
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, L, C)
    x = bert_output_ln(ctx, hidden, sd, pfx + "attention.output.")
    if enc_hidden is not None and (pfx + "crossattention.self.query.weight") in sd:
        c2, _ = bert_attention_core(x, enc_hidden, cross_mask, sd, pfx + "crossattention.self.", num_heads)
        x = bert_output_ln(c2, x, sd, pfx + "crossattention.output.")
    inter = F.gelu(F.linear(x, sd[pfx + "intermediate.dense.weight"], sd[pfx + "intermediate.dense.bias"]))
    return bert_output_ln(inter, x, sd, pfx + "output."), (k, v)


def bert_decoder(sd, pfx, num_heads, num_layers, input_ids, attention_mask=None, enc_hidden=None, enc_mask=None,
                 position_ids=None, past=None, history=None):
    """BertModel.forward with is_decoder=True (models/xbert.py:1075-1220, eval): embeddings at positions offset by the
    cache length, causal/3-D self mask, every layer through bert_layer_cached.
    Returns (last hidden, presents per layer, hidden states incl. the embedding output)."""
    B, L = input_ids.shape
    n_past = past[0][0].shape[2] if past is not None else 0
    if position_ids is None:
        position_ids = torch.arange(n_past, n_past + L)[None].expand(B, -1)
    e = pfx + "embeddings."
    h = sd[e + "word_embeddings.weight"][input_ids] + sd[e + "token_type_embeddings.weight"][0] \
        + sd[e + "position_embeddings.weight"][position_ids]
    h = F.layer_norm(h, (h.shape[-1],), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], 1e-12)
    if attention_mask is None:
        attention_mask = torch.ones(B, L + n_past)
    self_mask = decoder_self_mask(attention_mask, L)
    cross_mask = None
    if enc_hidden is not None:
        cross_mask = extended_cross_mask(enc_mask if enc_mask is not None else torch.ones(enc_hidden.shape[:2]), h.dtype)
    presents, hiddens = [], [h]
    for i in range(num_layers):
        h, kv = bert_layer_cached(h, self_mask, sd, "%sencoder.layer.%d." % (pfx, i), num_heads, enc_hidden, cross_mask,
                                  history=history[i] if history is not None else None,
                                  past=past[i] if past is not None else None)
        presents.append(kv)
        hiddens.append(h)
    return h, presents, hiddens


def lm_loss(logits, labels, label_smoothing=0.0, reduction="mean"):
    """Shifted next-token loss of BertLMHeadModel (models/xbert.py:1369-1383) with LabelSmoothSoftmaxCEV1
    (:1223-1263): smoothed target = (1-s) on the label, s/C elsewhere; ignore_index -100."""
    B, L, V = logits.shape
    lg = logits[:, :-1].reshape(-1, V)
    lb = labels[:, 1:].reshape(-1)
    if label_smoothing > 0:
        logp = lg.float().log_softmax(dim=1)
        ignore = lb.eq(-100)
        t = torch.full_like(logp, label_smoothing / V)
        t.scatter_(1, lb.masked_fill(ignore, 0)[:, None], 1.0 - label_smoothing)
        loss = -(logp * t).sum(dim=1)
        loss[ignore] = 0
        if reduction == "mean":
            return loss.sum() / ignore.eq(0).sum()
        loss = loss.sum() if reduction == "sum" else loss
    else:
        loss = F.cross_entropy(lg, lb, reduction=reduction)
    return loss.view(B, -1).sum(1) if reduction == "none" else loss


class Shapes:
    """Model shape of a config (base: 12 vision blocks / 12 heads, 12 text + 6 fusion layers)."""

    def __init__(self, vision_depth=12, vision_heads=12, text_heads=12, fusion_layer=12, num_layers=18):
        self.vision_depth, self.vision_heads, self.text_heads = vision_depth, vision_heads, text_heads
        self.fusion_layer, self.num_layers = fusion_layer, num_layers


def pretrain_forward(sd, shp, image, text_ids, text_atts, text_ids_masked, masked_pos, masked_ids, image_neg_idx,
                     text_neg_idx, image_atts=None, idx_to_group_img=None, target_bbox=None, is_image=None,
                     ret_bbox_loss=False, train=False, out=None):
    """XVLM.forward_multimodal (models/model_pretrain.py:30-65) with the hard-negative indices given
    (the reference draws them with torch.multinomial, models/xvlm.py:847-855).  Returns the dict of
    losses; intermediate tensors are stored into `out` if given."""
    vp, bp = "vision_encoder.", "text_encoder.bert."
    if ret_bbox_loss:
        image_embeds, full = vision_forward(image, sd, vp, shp.vision_depth, shp.vision_heads, idx_to_group_img, image_atts)
        image_embeds_fullatts = full[idx_to_group_img]
    else:
        image_embeds = vision_forward(image, sd, vp, shp.vision_depth, shp.vision_heads)
        image_atts = torch.ones(image_embeds.shape[:2], dtype=torch.long)
    bert = dict(sd=sd, pfx=bp, num_heads=shp.text_heads, fusion_layer=shp.fusion_layer, num_layers=shp.num_layers,
                train=train)
    text_embeds = bert_model(input_ids=text_ids, attention_mask=text_atts, mode="text", **bert)
    image_feat, text_feat = get_features(image_embeds, text_embeds, sd)
    temp = sd["temp"]
    loss = {"loss_itc": contrastive_loss(image_feat, text_feat, temp)}
    # ITM (models/xvlm.py:859-899)
    bs = image_feat.shape[0]
    img_neg, att_neg = image_embeds[image_neg_idx], image_atts[image_neg_idx]
    txt_neg, tatt_neg = text_embeds[text_neg_idx], text_atts[text_neg_idx]
    cross_pos = bert_model(encoder_embeds=text_embeds, attention_mask=text_atts, enc_hidden=image_embeds,
                           enc_mask=image_atts, mode="fusion", **bert)[:, 0]
    cross_neg = bert_model(encoder_embeds=torch.cat([text_embeds, txt_neg]), attention_mask=torch.cat([text_atts, tatt_neg]),
                           enc_hidden=torch.cat([img_neg, image_embeds]), enc_mask=torch.cat([att_neg, image_atts]),
                           mode="fusion", **bert)[:, 0]
    itm_logits = mlp_head(torch.cat([cross_pos, cross_neg]), sd, "itm_head.")
    itm_labels = torch.cat([torch.ones(bs, dtype=torch.long), torch.zeros(2 * bs, dtype=torch.long)])
    loss["loss_itm"] = F.cross_entropy(itm_logits, itm_labels)
    # MLM (models/xvlm.py:901-908, models/xbert.py:1591-1673)
    seq = bert_model(input_ids=text_ids_masked, attention_mask=text_atts, enc_hidden=image_embeds, enc_mask=image_atts,
                     mode="multi_modal", **bert)
    mlm_logits = mlm_head(gather_by_pos(seq, masked_pos), sd, "text_encoder.cls.predictions.")
    loss["loss_mlm"] = F.cross_entropy(mlm_logits.view(-1, mlm_logits.shape[-1]), masked_ids.view(-1))
    if ret_bbox_loss:
        # predict_bbox: fusion over the FULL-attention image embeds (models/xvlm.py:910-925)
        cls = bert_model(encoder_embeds=text_embeds, attention_mask=text_atts, enc_hidden=image_embeds_fullatts,
                         enc_mask=torch.ones(image_embeds_fullatts.shape[:2]), mode="fusion", **bert)[:, 0]
        coord = mlp_head(cls, sd, "bbox_head.").sigmoid()
        loss["loss_bbox"], loss["loss_giou"] = bbox_loss(coord, target_bbox, is_image)
        if out is not None:
            out["bbox_coord"] = coord
    if out is not None:
        out.update(image_embeds=image_embeds, text_embeds=text_embeds, image_feat=image_feat, text_feat=text_feat,
                   cross_pos=cross_pos, cross_neg=cross_neg, itm_logits=itm_logits, mlm_logits=mlm_logits)
    return loss


# ------------------------------------------------------------------------------------------------
# retrieval scoring (Retrieval.py:71-157) and video input (models/xvlm.py:615-661)
# ------------------------------------------------------------------------------------------------
def video_forward(frames, sd, shp, frame_pos=None):
    """video_encoding == 'avgpool': frames [B, F, 3, H, W] -> every frame through the vision encoder, + per-frame
    offset, mean over F (models/xvlm.py:615-645).  Returns [B, 197, D]."""
    B, n_f = frames.shape[:2]
    emb = vision_forward(frames.reshape(B * n_f, *frames.shape[2:]), sd, "vision_encoder.", shp.vision_depth, shp.vision_heads)
    emb = emb.view(B, n_f, emb.shape[1], emb.shape[2])
    if frame_pos is not None:
        emb = emb + frame_pos
    return emb.mean(dim=1)


def retrieval_scores(sd, shp, image_embeds, text_embeds, text_atts, sims_matrix, k_test, topk_i2t=None, topk_t2i=None):
    """The two re-ranking loops of Retrieval.py:118-151, one fusion call of k_test sequences per image / caption.
    `topk_*` override the candidate lists (to score exactly the pairs another implementation picked)."""
    bert = dict(sd=sd, pfx="text_encoder.bert.", num_heads=shp.text_heads, fusion_layer=shp.fusion_layer,
                num_layers=shp.num_layers)
    n_img, n_txt = sims_matrix.shape
    s_i2t = torch.full((n_img, n_txt), -100.0)
    s_t2i = torch.full((n_txt, n_img), -100.0)
    for i in range(n_img):
        idx = topk_i2t[i] if topk_i2t is not None else sims_matrix[i].topk(k=min(k_test, n_txt), dim=0).indices
        enc = image_embeds[i].repeat(len(idx), 1, 1)
        out = bert_model(encoder_embeds=text_embeds[idx], attention_mask=text_atts[idx], enc_hidden=enc,
                         enc_mask=torch.ones(enc.shape[:2], dtype=torch.long), mode="fusion", **bert)
        s_i2t[i, idx] = mlp_head(out[:, 0], sd, "itm_head.")[:, 1]
    for j in range(n_txt):
        idx = topk_t2i[j] if topk_t2i is not None else sims_matrix[:, j].topk(k=min(k_test, n_img), dim=0).indices
        enc = image_embeds[idx]
        out = bert_model(encoder_embeds=text_embeds[j].repeat(len(idx), 1, 1), attention_mask=text_atts[j].repeat(len(idx), 1),
                         enc_hidden=enc, enc_mask=torch.ones(enc.shape[:2], dtype=torch.long), mode="fusion", **bert)
        s_t2i[j, idx] = mlp_head(out[:, 0], sd, "itm_head.")[:, 1]
    return s_i2t, s_t2i
