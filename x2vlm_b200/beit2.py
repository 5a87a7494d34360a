"""Drop-in replacement of the reference's models/beit2.py hot path (BEiT-2 vision encoder).

Same constructor signatures, attribute names, forward signature/returns and state_dict keys as the
reference (`VisionTransformer`, `Block`, `Attention`, `Mlp`, `PatchEmbed`, `beit_base_patch16`,
`beit_large_patch16`; models/beit2.py:51-470) so models/xvlm.py's
`build_vision_encoder` (:246-283) constructs and calls it unchanged — but every block runs as one
autograd node of hand-written sm_100a kernels (x2vlm_b200.functional.beit_block).

The CUDA extension is mandatory: forward on a CPU tensor raises (no eager fallback).
"""
import math
from functools import partial

import torch
import torch.nn as nn

from . import functional as XF
from .params import GappedBias, Shadow


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm 0.4.9 semantics); only carries the rate — the scaling is fused
    into the block's GEMM epilogues."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def sample_scale(self, batch, device):
        """[B] fp32: floor(keep + U[0,1)) / keep, or None when inactive."""
        if not self.training or not self.drop_prob:
            return None
        keep = 1.0 - self.drop_prob
        return torch.floor(keep + torch.rand(batch, device=device, dtype=torch.float32)) / keep

    def forward(self, x):
        s = self.sample_scale(x.shape[0], x.device)
        return x if s is None else x * s.view(-1, *([1] * (x.dim() - 1)))

    def extra_repr(self):
        return "p={}".format(self.drop_prob)


class Mlp(nn.Module):
    """fc2(GELU(fc1(x))) — parameters only; executed inside the fused block (models/beit2.py:51-68)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    """Parameter holder with the reference's names (models/beit2.py:71-123)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., window_size=None,
                 attn_head_dim=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if attn_head_dim is not None:
            head_dim = attn_head_dim
        if head_dim != 64:
            raise NotImplementedError("x2k attention kernels are built for head_dim 64 (got %d)" % head_dim)
        if attn_drop or proj_drop:
            raise NotImplementedError("BEiT attn_drop / proj_drop are 0 in every X2-VLM config (models/xvlm.py:259-261)")
        all_head_dim = head_dim * num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, all_head_dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(all_head_dim))
            self.v_bias = nn.Parameter(torch.zeros(all_head_dim))
        else:
            self.q_bias = None
            self.v_bias = None
        if window_size:
            self.window_size = window_size
            wh, ww = window_size
            self.num_relative_distance = (2 * wh - 1) * (2 * ww - 1) + 3
            self.relative_position_bias_table = nn.Parameter(torch.zeros(self.num_relative_distance, num_heads))
            # pair-wise relative position index (cls<->token / token<->cls / cls<->cls use the last 3 rows)
            ch, cw = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
            coords = torch.stack([ch.reshape(-1), cw.reshape(-1)])
            rel = coords[:, :, None] - coords[:, None, :]
            index = torch.zeros(wh * ww + 1, wh * ww + 1, dtype=torch.int64)
            index[1:, 1:] = (rel[0] + wh - 1) * (2 * ww - 1) + (rel[1] + ww - 1)
            index[0, :] = self.num_relative_distance - 3
            index[:, 0] = self.num_relative_distance - 2
            index[0, 0] = self.num_relative_distance - 1
            self.register_buffer("relative_position_index", index)
        else:
            self.window_size = None
            self.relative_position_bias_table = None
            self.relative_position_index = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(all_head_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, window_size=None, attn_head_dim=None):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, window_size=window_size, attn_head_dim=attn_head_dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if init_values is not None and init_values > 0:
            self.gamma_1 = nn.Parameter(init_values * torch.ones(dim), requires_grad=True)
            self.gamma_2 = nn.Parameter(init_values * torch.ones(dim), requires_grad=True)
        else:
            self.gamma_1, self.gamma_2 = None, None
        # bf16 shadows / gradient sinks of the four GEMM weights
        self._x2k = {"qkv": Shadow(self.attn.qkv.weight), "proj": Shadow(self.attn.proj.weight),
                     "fc1": Shadow(self.mlp.fc1.weight), "fc2": Shadow(self.mlp.fc2.weight)}
        if self.attn.q_bias is not None:
            self._x2k["bqv"] = GappedBias(self.attn.q_bias, self.attn.v_bias)
        self._x2k_shadows = list(self._x2k.values())

    def forward(self, x, rel_pos_bias=None, return_attention=False, return_qkv=False, image_atts=None,
                output_attentions=None, drop_path_scales=None):
        if return_attention or return_qkv or rel_pos_bias is not None or image_atts is not None:
            # the reference never passes these on the hot path (VisionTransformer.forward, beit2.py:401-407)
            raise NotImplementedError("x2k Block: return_attention / return_qkv / rel_pos_bias / image_atts "
                                      "are outside the fused hot path")
        # attention maps are rebuilt on request only (knowledge distillation): the fused block never materialises them
        attn_prob = XF.beit_attention_map(x, self) if output_attentions else None
        dp1 = dp2 = None
        if drop_path_scales is not None:  # pre-sampled for the whole encoder in one shot (VisionTransformer.forward_blocks)
            dp1, dp2 = drop_path_scales
        elif isinstance(self.drop_path, DropPath):  # two independent per-sample draws per block (beit2.py:204-207)
            dp1 = self.drop_path.sample_scale(x.shape[0], x.device)
            dp2 = self.drop_path.sample_scale(x.shape[0], x.device)
        return XF.beit_block(x, self, dp1, dp2), attn_prob


class PatchEmbed(nn.Module):
    """Conv2d(k = s = patch) as an im2col GEMM on the tcgen05 kernel (models/beit2.py:212-232)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.patch_shape = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.img_size = img_size
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self._shadow = Shadow(self.proj.weight, K=in_chans * patch_size[0] * patch_size[1])
        self._x2k_shadows = [self._shadow]

    def forward(self, x, **kwargs):
        B, C, H, W = x.shape
        ph, pw = self.patch_size
        gh, gw = H // ph, W // pw
        # non-overlapping patches -> rows of a [B*gh*gw, C*ph*pw] matrix (same (c, i, j) order as the conv weight)
        cols = x.reshape(B, C, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * ph * pw)
        y = XF.linear(cols, self._shadow, self.proj.bias)
        return y.reshape(B, gh * gw, -1)


class VisionTransformer(nn.Module):
    """Same constructor and forward contract as models/beit2.py:274-436."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, init_values=None, use_abs_pos_emb=True, use_rel_pos_bias=False,
                 use_shared_rel_pos_bias=False, use_mean_pooling=True, init_scale=0.001, local_attn_depth=-1,
                 vision_num_hidden_layers=-1):
        super().__init__()
        self.local_attn_depth = -1
        if vision_num_hidden_layers > 0:
            depth = vision_num_hidden_layers
        self.depth = depth
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        if use_shared_rel_pos_bias or not use_mean_pooling or drop_rate:
            raise NotImplementedError("x2k VisionTransformer covers the X2-VLM configuration: per-block relative position "
                                      "bias, mean pooling, drop_rate 0 (models/xvlm.py:259-265)")
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim)) if use_abs_pos_emb else None
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.rel_pos_bias = None
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.use_rel_pos_bias = use_rel_pos_bias
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                  init_values=init_values, window_size=self.patch_embed.patch_shape if use_rel_pos_bias else None)
            for i in range(depth)])
        self.norm = nn.Identity()
        self.fc_norm = norm_layer(embed_dim)
        self.register_buffer("_x2k_dp_keep", 1.0 - torch.tensor(dpr, dtype=torch.float32).view(-1, 1, 1), persistent=False)
        self._x2k_dp_any = max(dpr) > 0
        if self.pos_embed is not None:
            _trunc_normal_(self.pos_embed, std=.02)
        _trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)
        self.fix_init_weight()

    def fix_init_weight(self):
        for layer_id, layer in enumerate(self.blocks):
            layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
            layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def forward_blocks(self, x, all_states=None, all_attentions=None):
        """patch embed -> cls cat -> blocks.  Returns (last block output [B, N, D] incl. the cls row, hidden states or None)."""
        x = self.patch_embed(x)
        batch_size = x.shape[0]
        x = torch.cat((self.cls_token.expand(batch_size, -1, -1), x), dim=1)
        if self.pos_embed is not None:
            if x.shape[1] != self.pos_embed.shape[1]:
                raise NotImplementedError("pos_embed interpolation (use_abs_pos_emb is False in X2-VLM)")
            x = x + self.pos_embed
        # stochastic depth: the 2 x depth per-sample masks of this pass in three kernels instead of six per block
        dps = None
        if self.training and self._x2k_dp_any:
            keep = self._x2k_dp_keep  # [depth, 1, 1] on the module's device (non-persistent buffer)
            dps = torch.floor(keep + torch.rand(len(self.blocks), 2, batch_size, device=x.device)) / keep
        for i, blk in enumerate(self.blocks):
            if all_states is not None:
                all_states = all_states + (x,)
            blk_dps = (dps[i, 0], dps[i, 1]) if (dps is not None and isinstance(blk.drop_path, DropPath)) else None
            x, attn = blk(x, output_attentions=all_attentions is not None, drop_path_scales=blk_dps)
            if all_attentions is not None:
                all_attentions.append(attn)
        return x, all_states

    def tail(self, x, idx_to_group_img=None, image_atts=None):
        """models/beit2.py:409-436 as one kernel (x2k_pool_tail): the cls output is dropped, fc_norm over the patch tokens,
        their mean becomes token 0; with idx_to_group_img / image_atts the per-region gather + mask-weighted mean."""
        return XF.pool_tail(x.float(), self.fc_norm.weight, self.fc_norm.bias, self.fc_norm.eps, idx_to_group_img, image_atts)

    def forward_features(self, x, all_states=None, all_attentions=None):
        """Returns (normalised patch tokens [B,P,D], their mean [B,1,D], hidden states tuple or None)."""
        x, all_states = self.forward_blocks(x, all_states, all_attentions)
        full = self.tail(x)
        return full[:, 1:], full[:, :1], all_states

    @staticmethod
    def region_pool(x, idx_to_group_img, image_atts):
        """Per-region gather + mask-weighted mean of the (already normalised) patch tokens (beit2.py:430-434), in torch."""
        x_bs = x[idx_to_group_img]
        weights = image_atts[:, 1:].unsqueeze(2).to(x.dtype)
        x_bs_cls = (weights * x_bs).sum(dim=1, keepdim=True) / weights.sum(dim=1, keepdim=True)
        return torch.cat([x_bs_cls, x_bs], dim=1)

    def forward(self, x, idx_to_group_img=None, image_atts=None, output_attentions=None, output_hidden_states=None):
        assert output_attentions == output_hidden_states
        attns = [] if output_attentions else None
        x, all_states = self.forward_blocks(x, () if output_hidden_states else None, attns)
        full = self.tail(x)
        if idx_to_group_img is None:
            if output_hidden_states:
                all_states = all_states + (full,)
                assert len(all_states) == len(attns) + 1
                return {'last_hidden_state': full, 'hidden_states': all_states, 'attentions': tuple(attns)}
            return full
        if output_hidden_states:
            raise NotImplementedError("not implemented KD for BBox Loss")
        return self.tail(x, idx_to_group_img, image_atts), full


def beit_base_patch16(img_size, **kwargs):
    return VisionTransformer(img_size=img_size, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def beit_large_patch16(img_size, **kwargs):
    return VisionTransformer(img_size=img_size, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


# ----------------------------------------------------------------------------------------------
# checkpoint helpers (host logic; models/beit2.py:473-754)
# ----------------------------------------------------------------------------------------------
def _resize_rel_pos_bias_table(rel_pos_bias, src_size, dst_size, num_extra_tokens):
    """Relative-position bias table [(2·src-1)² + extra, H] -> [(2·dst-1)² + extra, H] (models/beit2.py:514-575):
    source offsets are placed on a geometric progression (denser near 0), every head's (2·src-1)² surface is fitted
    with a bicubic spline and sampled on the integer offsets of the target window; the `extra` rows (cls-to-token,
    token-to-cls, cls-to-cls) are copied.  The reference calls scipy.interpolate.interp2d(kind='cubic'), which SciPy
    removed in 1.14; RectBivariateSpline(kx=ky=3) on the same rectilinear grid is SciPy's documented replacement
    (interp2d(x, y, z)(xn, yn) == RectBivariateSpline(x, y, z.T)(xn, yn).T)."""
    import numpy as np
    from scipy import interpolate
    extra_tokens = rel_pos_bias[-num_extra_tokens:, :]
    table = rel_pos_bias[:-num_extra_tokens, :]
    num_heads = table.shape[1]

    def geometric_progression(a, r, n):
        return a * (1.0 - r ** n) / (1.0 - r)

    left, right = 1.01, 1.5
    while right - left > 1e-6:
        q = (left + right) / 2.0
        if geometric_progression(1, q, src_size // 2) > dst_size // 2:
            right = q
        else:
            left = q
    dis, cur = [], 1
    for i in range(src_size // 2):
        dis.append(cur)
        cur += q ** (i + 1)
    x = [-d for d in reversed(dis)] + [0] + dis
    t = dst_size // 2.0
    dx = np.arange(-t, t + 0.1, 1.0)
    out = []
    for h in range(num_heads):
        z = table[:, h].view(src_size, src_size).float().cpu().numpy()
        spline = interpolate.RectBivariateSpline(x, x, z.T, kx=3, ky=3)
        out.append(torch.tensor(spline(dx, dx).T, dtype=torch.float32).contiguous().view(-1, 1).to(rel_pos_bias.device))
    return torch.cat((torch.cat(out, dim=-1).to(rel_pos_bias.dtype), extra_tokens), dim=0)


def interpolate_pos_embed(model, checkpoint_model):
    """Adapt a vision state_dict (no prefix) to `model`'s resolution in place (models/beit2.py:664-754): drop the
    `relative_position_index` buffers (rebuilt by the model), resize every `relative_position_bias_table` whose window
    differs, bicubic-resize `pos_embed` if both sides have one.  Returns the dict."""
    for key in list(checkpoint_model.keys()):
        if "relative_position_index" in key:
            checkpoint_model.pop(key)
        if "relative_position_bias_table" in key:
            rel_pos_bias = checkpoint_model[key]
            src_num_pos, _ = rel_pos_bias.size()
            try:
                dst_num_pos, _ = model.state_dict()[key].size()
            except KeyError:
                print("Note that vision encoder does not have: ", key)
                continue
            dst_patch_shape = model.patch_embed.patch_shape
            if dst_patch_shape[0] != dst_patch_shape[1]:
                raise NotImplementedError()
            num_extra_tokens = dst_num_pos - (dst_patch_shape[0] * 2 - 1) * (dst_patch_shape[1] * 2 - 1)
            src_size = int((src_num_pos - num_extra_tokens) ** 0.5)
            dst_size = int((dst_num_pos - num_extra_tokens) ** 0.5)
            if src_size != dst_size:
                print("Position interpolate for %s from %dx%d to %dx%d" % (key, src_size, src_size, dst_size, dst_size))
                checkpoint_model[key] = _resize_rel_pos_bias_table(rel_pos_bias, src_size, dst_size, num_extra_tokens)
    if ('pos_embed' in checkpoint_model) and (model.pos_embed is not None):
        pos_embed_checkpoint = checkpoint_model['pos_embed']
        embedding_size = pos_embed_checkpoint.shape[-1]
        num_patches = model.patch_embed.num_patches
        num_extra_tokens = model.pos_embed.shape[-2] - num_patches
        orig_size = int((pos_embed_checkpoint.shape[-2] - num_extra_tokens) ** 0.5)
        new_size = int(num_patches ** 0.5)
        if orig_size != new_size:
            print("Position interpolate from %dx%d to %dx%d" % (orig_size, orig_size, new_size, new_size))
            extra_tokens = pos_embed_checkpoint[:, :num_extra_tokens]
            pos_tokens = pos_embed_checkpoint[:, num_extra_tokens:]
            pos_tokens = pos_tokens.reshape(-1, orig_size, orig_size, embedding_size).permute(0, 3, 1, 2)
            pos_tokens = torch.nn.functional.interpolate(pos_tokens, size=(new_size, new_size), mode='bicubic',
                                                         align_corners=False)
            pos_tokens = pos_tokens.permute(0, 2, 3, 1).flatten(1, 2)
            checkpoint_model['pos_embed'] = torch.cat((extra_tokens, pos_tokens), dim=1)
    return checkpoint_model


def load_state_dict(model, state_dict, prefix='', ignore_missing="relative_position_index"):
    """Non-strict recursive load with the reference's reporting (models/beit2.py:617-661): missing keys that contain
    one of the `|`-separated `ignore_missing` patterns are expected.  Returns (missing, unexpected, ignored)."""
    missing_keys, unexpected_keys, error_msgs = [], [], []
    metadata = getattr(state_dict, '_metadata', None)
    state_dict = state_dict.copy()
    if metadata is not None:
        state_dict._metadata = metadata

    def load(module, prefix=''):
        local_metadata = {} if metadata is None else metadata.get(prefix[:-1], {})
        module._load_from_state_dict(state_dict, prefix, local_metadata, True, missing_keys, unexpected_keys, error_msgs)
        for name, child in module._modules.items():
            if child is not None:
                load(child, prefix + name + '.')

    load(model, prefix=prefix)
    warn, ignored = [], []
    for key in missing_keys:
        (ignored if any(pat in key for pat in ignore_missing.split('|')) else warn).append(key)
    if warn:
        print("Weights of {} not initialized from pretrained model: {}".format(model.__class__.__name__, warn))
    if unexpected_keys:
        print("Weights from pretrained model not used in {}: {}".format(model.__class__.__name__, unexpected_keys))
    if ignored:
        print("Ignored weights of {} not initialized from pretrained model: {}".format(model.__class__.__name__, ignored))
    if error_msgs:
        print('\n'.join(error_msgs))
    # (parameters are overwritten in place: their version counters change, so the bf16 shadows re-cast on next use)
    return warn, unexpected_keys, ignored


def load_pretrained_beit2(model, ckpt_rpath):
    """Load a BEiT-2 checkpoint file into the vision encoder (models/beit2.py:473-614): unwrap 'model' / 'module',
    drop the classification head, expand a shared rel-pos bias to every block, adapt tables / pos_embed to the
    model's resolution, then load non-strictly."""
    print("Load BEIT-V2 ckpt from %s" % ckpt_rpath)
    checkpoint = torch.load(ckpt_rpath, map_location='cpu')
    checkpoint_model = None
    for model_key in 'model|module'.split('|'):
        if model_key in checkpoint:
            checkpoint_model = checkpoint[model_key]
            print("Load state_dict by model_key = %s" % model_key)
            break
    if checkpoint_model is None:
        checkpoint_model = checkpoint
    for k in ['head.weight', 'head.bias']:
        del checkpoint_model[k]
    if getattr(model, 'use_rel_pos_bias', False) and "rel_pos_bias.relative_position_bias_table" in checkpoint_model:
        print("Expand the shared relative position embedding to each transformer block. ")
        rel_pos_bias = checkpoint_model.pop("rel_pos_bias.relative_position_bias_table")
        for i in range(model.get_num_layers()):
            checkpoint_model["blocks.%d.attn.relative_position_bias_table" % i] = rel_pos_bias.clone()
    interpolate_pos_embed(model, checkpoint_model)
    return load_state_dict(model, checkpoint_model, prefix='')
