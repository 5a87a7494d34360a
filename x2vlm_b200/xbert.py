"""Drop-in replacement of the reference's models/xbert.py hot path (BERT text encoder, cross-modal
fusion layers, MLM head).

Same class names, constructor arguments, attribute paths, forward keyword arguments and state_dict
keys as the reference (`BertConfig`, `BertModel`, `BertForMaskedLM`, `BertOnlyMLMHead`,
`BertLayer`, …; models/xbert.py:169-1687), so models/xvlm.py's
`build_text_encoder` (:286-316) and `XVLMBase.get_text_embeds / get_cross_embeds / get_mlm_loss`
construct and call it unchanged.  Each `BertLayer` executes as ONE autograd node of hand-written
sm_100a kernels (x2vlm_b200.functional.bert_layer): packed QKV GEMM, tcgen05 attention with masks and
in-kernel Philox dropout, dense+dropout+residual epilogues, LayerNorm kernels.

One extension beyond the reference API: `encoder_kv_index` (int32 [B]) lets several text sequences
attend to the SAME image's keys/values — the image K/V projection of a fusion layer is then computed
once per image instead of once per (text, image) pair.

Generation paths (`history_states` of the captioning beam search, model_generation.py:180-188; HF-style
`past_key_values` / `use_cache` of `BertLMHeadModel`, xbert.py:355-359) run forward-only on the same
kernels (x2vlm_b200.functional.bert_layer_decode) and refuse to record an autograd graph.

`output_attentions` / `save_attention` are served on request: a streaming kernel (x2k_attn_probs) rebuilds the
pre-dropout maps (and, for save_attention, their gradients) from Q, K and the fused forward's log-sum-exp.
Not built (raise NotImplementedError): head pruning / head_mask, split_lengths.
xbert's per-position DropPath (text_drop_path_rate / cross_drop_path_rate, refcoco_grounding_large.yaml) is folded into
the dense GEMM epilogues as a per-row scale.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F
from transformers import BertConfig  # noqa: F401  (re-exported, models/xvlm.py:25 imports it from here)
from transformers.modeling_outputs import (BaseModelOutputWithPastAndCrossAttentions,
                                           BaseModelOutputWithPoolingAndCrossAttentions, CausalLMOutputWithCrossAttentions,
                                           MaskedLMOutput)

from . import functional as XF
from . import ops
from .params import Shadow


class BertEmbeddings(nn.Module):
    """word + position + token_type embeddings -> LayerNorm -> dropout (models/xbert.py:169-216)."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.position_embedding_type = getattr(config, "position_embedding_type", "absolute")
        self.config = config

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None, past_key_values_length=0):
        input_shape = input_ids.size() if input_ids is not None else inputs_embeds.size()[:-1]
        seq_length = input_shape[1]
        explicit_pos = position_ids is not None
        if position_ids is None:
            position_ids = self.position_ids[:, past_key_values_length: seq_length + past_key_values_length]
        if inputs_embeds is None and self.position_embedding_type == "absolute" and input_ids.is_cuda:
            # one kernel: three gathers + LayerNorm + dropout (x2k_embed_ln_fwd); token_type_ids None = all zeros
            return XF.embed_ln(input_ids, token_type_ids, position_ids if explicit_pos else None, past_key_values_length,
                               self.word_embeddings.weight, self.position_embeddings.weight, self.token_type_embeddings.weight,
                               self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps, self.dropout.p, self.training)
        if token_type_ids is None:
            token_type_ids = torch.zeros(input_shape, dtype=torch.long, device=self.position_ids.device)
        if inputs_embeds is None:
            inputs_embeds = self.word_embeddings(input_ids)
        embeddings = inputs_embeds + self.token_type_embeddings(token_type_ids)
        if self.position_embedding_type == "absolute":
            embeddings = embeddings + self.position_embeddings(position_ids)
        embeddings = XF.layer_norm(embeddings, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps)
        return self.dropout(embeddings)


class BertSelfAttention(nn.Module):
    """Parameter holder (models/xbert.py:219-260); the math runs inside the fused layer."""

    def __init__(self, config, is_cross_attention):
        super().__init__()
        self.config = config
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("hidden size %d is not a multiple of the number of heads %d"
                             % (config.hidden_size, config.num_attention_heads))
        self.fp16 = getattr(config, 'fp16', False)
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        if self.attention_head_size != 64:
            raise NotImplementedError("x2k attention kernels are built for head_dim 64")
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.key = nn.Linear(kv_in, self.all_head_size)
        self.value = nn.Linear(kv_in, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self.position_embedding_type = getattr(config, "position_embedding_type", "absolute")
        if self.position_embedding_type != "absolute":
            raise NotImplementedError("relative position embeddings (the reference raises too, xbert.py:381-382)")
        self.save_attention = False

    # Grad-CAM hooks (models/xbert.py:248-260): with save_attention = True on a cross-attention module the fused layer
    # stores the pre-dropout map [B, H, L, Nk] after forward and its gradient after backward
    def save_attn_gradients(self, attn_gradients):
        self.attn_gradients = attn_gradients

    def get_attn_gradients(self):
        return self.attn_gradients

    def save_attention_map(self, attention_map):
        self.attention_map = attention_map

    def get_attention_map(self):
        return self.attention_map


class DropPath(nn.Module):
    """xbert's stochastic depth (models/xbert.py:518-548): unlike timm's per-sample version the random mask has shape
    (1, L, 1) — one draw per sequence POSITION, shared by the whole batch.  The fused layer reads `drop_prob` and
    folds the mask into the dense GEMM's epilogue (per-row scale); `forward` is the plain-torch definition."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep_prob = 1 - self.drop_prob
        mask = torch.floor(keep_prob + torch.rand((1, x.shape[1], 1), dtype=x.dtype, device=x.device))
        return x.div(keep_prob) * mask

    def extra_repr(self):
        return "p={}".format(self.drop_prob)


class BertSelfOutput(nn.Module):
    def __init__(self, config, drop_path_rate=0.0):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 1e-3 else nn.Identity()


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False, drop_path_rate=0.0):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config, drop_path_rate=drop_path_rate)
        self.pruned_heads = set()


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        if config.hidden_act not in ("gelu", F.gelu):
            raise NotImplementedError("fused FFN epilogue implements exact (erf) GELU only, got %r" % (config.hidden_act,))


class BertOutput(nn.Module):
    def __init__(self, config, drop_path_rate=0.0):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 1e-3 else nn.Identity()


class BertLayer(nn.Module):
    """self-attn -> (cross-attn when layer_num >= fusion_layer and encoder states are given) -> FFN
    (models/xbert.py:551-625)."""

    def __init__(self, config, layer_num, drop_path_rate=0.0):
        super().__init__()
        self.config = config
        self.chunk_size_feed_forward = getattr(config, "chunk_size_feed_forward", 0)
        self.seq_len_dim = 1
        self.attention = BertAttention(config, drop_path_rate=drop_path_rate)
        self.has_cross_attention = (layer_num >= config.fusion_layer)
        if self.has_cross_attention:
            self.layer_num = layer_num
            self.crossattention = BertAttention(config, is_cross_attention=True, drop_path_rate=drop_path_rate)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config, drop_path_rate=drop_path_rate)
        at = self.attention
        D = config.hidden_size
        sh = {"qkv": Shadow(at.self.query.weight, at.self.key.weight, at.self.value.weight),
              "bqkv": Shadow(at.self.query.bias, at.self.key.bias, at.self.value.bias, K=D),
              "o": Shadow(at.output.dense.weight)}
        if self.has_cross_attention:
            ca = self.crossattention
            sh.update(qc=Shadow(ca.self.query.weight), kvc=Shadow(ca.self.key.weight, ca.self.value.weight),
                      bkvc=Shadow(ca.self.key.bias, ca.self.value.bias, K=D), oc=Shadow(ca.output.dense.weight))
        sh.update(i=Shadow(self.intermediate.dense.weight), out=Shadow(self.output.dense.weight))
        self._x2k = sh
        self._x2k_shadows = list(sh.values())

    def fused(self, x, xb, cfg, enc=None, encb=None):
        return XF.bert_layer(x, xb, self, cfg, enc, encb)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, split_lengths=None,
                history_states=None):
        """Reference-shaped entry point: extended additive masks in, tuple out (layer output, present key/value)."""
        if head_mask is not None or split_lengths:
            raise NotImplementedError("x2k BertLayer: head_mask / split_lengths")
        B, L = hidden_states.shape[:2]
        decode = past_key_value is not None or history_states is not None
        n_past = (past_key_value[0].shape[2] if past_key_value is not None else 0) + \
                 (history_states.shape[1] if history_states is not None else 0)
        cfg = _layer_cfg(self.config, self.training and not decode, attention_mask, encoder_attention_mask, None, B, L,
                         encoder_hidden_states.shape[1] if encoder_hidden_states is not None else 0, hidden_states.device,
                         Lk_self=L + n_past)
        if encoder_hidden_states is not None:
            cfg["n_kv"] = encoder_hidden_states.shape[0]
            _group_cross(cfg, B, L, encoder_hidden_states.shape[1], hidden_states.device)
        if decode:
            _no_grad_only("past_key_value / history_states")
            out = XF.bert_layer_decode(hidden_states, self, cfg, encoder_hidden_states, history=history_states,
                                       past_kv=past_key_value[:2] if past_key_value else None, return_probs=bool(output_attentions))
            return (out[0], *[p for p in out[2] if p is not None], out[1]) if output_attentions else (out[0], out[1])
        y, _ = self.fused(hidden_states, None, cfg, encoder_hidden_states)
        if output_attentions:  # (layer output, self map, [cross map], present) like models/xbert.py:576-625
            maps = XF.bert_layer_decode(hidden_states.detach(), self, dict(cfg, train=False), encoder_hidden_states, return_probs=True)[2]
            return (y, *[p for p in maps if p is not None], None)
        return (y, None)


def _no_grad_only(what):
    if torch.is_grad_enabled():
        raise NotImplementedError("x2k xbert: %s is a forward-only (generation) path; call it under torch.no_grad()" % what)


def _pad_mask(ext_mask, B, Lq, Lk, device):
    """Extended additive mask [B,1,1,Lk] / [B,1,Lq,Lk] (or None) -> (fp32 [B,(Lq,)pad32(Lk)] contiguous, is_3d)."""
    if ext_mask is None:
        return None, False
    m = ext_mask.to(device=device, dtype=torch.float32)
    if m.dim() == 4:
        m = m[:, 0]
    if m.dim() != 3:
        raise ValueError("attention mask must be extended to [B,1,1|Lq,Lk], got %s" % (tuple(ext_mask.shape),))
    if m.shape[0] != B:
        m = m.expand(B, -1, -1)
    per_query = m.shape[1] != 1
    ld = ops.pad32(Lk)
    out = torch.zeros(B, m.shape[1], ld, dtype=torch.float32, device=device)
    out[:, :, :Lk] = m
    return (out if per_query else out[:, 0].contiguous()), per_query


def _layer_cfg(config, training, ext_self_mask, ext_cross_mask, kv_index, B, L, Nk, device, Lk_self=None):
    self_mask, self_3d = _pad_mask(ext_self_mask, B, L, Lk_self if Lk_self is not None else L, device)
    cross_mask, cross_3d = _pad_mask(ext_cross_mask, B, L, Nk, device) if Nk else (None, False)
    return dict(self_mask=self_mask, self_mask_3d=self_3d, cross_mask=cross_mask, cross_mask_3d=cross_3d, kv_index=kv_index,
                n_kv=0, kv_groups=None,
                train=bool(training), p_hidden=float(config.hidden_dropout_prob),
                p_attn=float(config.attention_probs_dropout_prob), eps=float(config.layer_norm_eps))


def _group_cross(cfg, B, L, Nk, device):
    """Short captions: group the query sequences by the K/V source they attend to (one table per encoder call,
    shared by every fusion layer) so the grouped cross-attention kernels can stack them in one MMA tile."""
    kv = cfg["kv_index"]
    if kv is None:
        kv = torch.arange(B, dtype=torch.int32, device=device)
    cfg["kv_groups"] = ops.attn_group_table(kv, cfg["n_kv"], L, Nk)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        text_dpr = getattr(config, 'text_drop_path_rate', 0.0)
        cross_dpr = getattr(config, 'cross_drop_path_rate', 0.0)
        if text_dpr > 0:
            assert cross_dpr > 0
            config.hidden_dropout_prob = 0.0  # the reference zeroes it when drop path is on (xbert.py:637-640)
        n_text, n_cross = config.fusion_layer, config.num_hidden_layers - config.fusion_layer
        dpr = [x.item() for x in torch.linspace(0, text_dpr, n_text)] + [x.item() for x in torch.linspace(0, cross_dpr, n_cross)]
        assert len(dpr) == config.num_hidden_layers
        self.layer = nn.ModuleList([BertLayer(config, i, dpr[i]) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_values=None, use_cache=None, output_attentions=False,
                output_hidden_states=False, return_dict=True, mode='multi_modal', split_lengths=None, history_states=None,
                encoder_kv_index=None):
        if split_lengths:
            raise NotImplementedError("x2k BertEncoder: split_lengths")
        if head_mask is not None and any(h is not None for h in head_mask):
            raise NotImplementedError("x2k BertEncoder: head_mask")
        if mode == 'text':
            start_layer, output_layer = 0, self.config.fusion_layer
        elif mode == 'fusion':
            start_layer, output_layer = self.config.fusion_layer, self.config.num_hidden_layers
        elif mode == 'multi_modal':
            start_layer, output_layer = 0, self.config.num_hidden_layers
        else:
            raise ValueError(f"mode {mode} is not supported")
        B, L, _ = hidden_states.shape
        enc = encoder_hidden_states
        if isinstance(enc, list):
            raise ValueError("no this case since we do not use ALBEF-NLVR anymore")
        Nk = enc.shape[1] if enc is not None else 0
        decode = bool(use_cache) or past_key_values is not None or history_states is not None
        n_past = 0
        if past_key_values is not None:
            n_past = past_key_values[start_layer][0].shape[2]
        elif history_states is not None:
            assert isinstance(history_states, list) and len(history_states) == (output_layer - start_layer + 1)
            n_past = history_states[0].shape[1]
        cfg = _layer_cfg(self.config, self.training and not decode, attention_mask, encoder_attention_mask, encoder_kv_index,
                         B, L, Nk, hidden_states.device, Lk_self=L + n_past)
        encb = None
        if enc is not None:
            enc = enc.float().contiguous()
            if encoder_kv_index is None and enc.shape[0] != B:
                raise ValueError("encoder_hidden_states batch %d != %d (pass encoder_kv_index to share K/V)" % (enc.shape[0], B))
            cfg["n_kv"] = enc.shape[0]
            _group_cross(cfg, B, L, Nk, hidden_states.device)
            if output_layer > self.config.fusion_layer:
                encb = ops.to_bf16(enc)  # one bf16 copy feeds the K/V projection of every fusion layer
            # no intermediate state leaves this call: the image-state gradients of all fusion layers can share one buffer
            cfg["fuse_d_enc"] = not output_hidden_states
        all_hidden_states = () if output_hidden_states else None
        # attention maps are rebuilt on request by a forward-only pass of the layer (the fused layer never holds them)
        all_self_attentions = () if output_attentions else None
        all_cross_attentions = () if (output_attentions and enc is not None) else None
        next_decoder_cache = () if use_cache else None
        x, xb = hidden_states.float().contiguous(), None
        if decode:
            _no_grad_only("use_cache / past_key_values / history_states")
        for i in range(start_layer, output_layer):
            if output_hidden_states:
                all_hidden_states = all_hidden_states + (x,)
            if decode:
                out = XF.bert_layer_decode(
                    x, self.layer[i], cfg, enc, encb,
                    history=history_states[i - start_layer] if history_states is not None else None,
                    past_kv=past_key_values[i][:2] if past_key_values is not None else None, return_probs=bool(output_attentions))
                x, present = out[0], out[1]
                maps = out[2] if output_attentions else None
                if use_cache:
                    next_decoder_cache += (present,)
            else:
                maps = XF.bert_layer_decode(x.detach(), self.layer[i], dict(cfg, train=False), enc, encb, return_probs=True)[2] \
                    if output_attentions else None
                x, xb = self.layer[i].fused(x, xb, cfg, enc, encb)
            if output_attentions:
                all_self_attentions = all_self_attentions + (maps[0],)
                if maps[1] is not None:
                    all_cross_attentions = all_cross_attentions + (maps[1],)
        if output_hidden_states:
            all_hidden_states = all_hidden_states + (x,)
        if not return_dict:
            return tuple(v for v in [x, next_decoder_cache, all_hidden_states, all_self_attentions, all_cross_attentions]
                         if v is not None)
        return BaseModelOutputWithPastAndCrossAttentions(last_hidden_state=x, past_key_values=next_decoder_cache,
                                                         hidden_states=all_hidden_states, attentions=all_self_attentions,
                                                         cross_attentions=all_cross_attentions)


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class BertPredictionHeadTransform(nn.Module):
    """dense + GELU + LayerNorm (models/xbert.py:785-802)."""

    def __init__(self, config):
        super().__init__()
        output_size = getattr(config, 'embedding_dim', config.hidden_size)
        self.dense = nn.Linear(config.hidden_size, output_size)
        self.LayerNorm = nn.LayerNorm(output_size, eps=config.layer_norm_eps)
        self._shadow = Shadow(self.dense.weight)
        self._x2k_shadows = [self._shadow]

    def forward(self, hidden_states):
        shp = hidden_states.shape
        h = XF.linear(hidden_states.reshape(-1, shp[-1]), self._shadow, self.dense.bias, gelu=True)
        h = XF.layer_norm(h, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps)
        return h.reshape(*shp[:-1], -1)


class BertLMPredictionHead(nn.Module):
    """transform + decoder tied to the word embeddings + output-only bias (models/xbert.py:805-823)."""

    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        hidden_size = getattr(config, 'embedding_dim', config.hidden_size)
        self.decoder = nn.Linear(hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias  # same Parameter under both state_dict keys, as in the reference
        self._shadow = None
        self._x2k_shadows = []

    def _x2k_refresh(self):
        self._decoder_shadow()

    def _decoder_shadow(self):
        if self._shadow is None or self._shadow.params[0] is not self.decoder.weight:
            self._shadow = Shadow(self.decoder.weight)
            self._x2k_shadows = [self._shadow]
        return self._shadow

    def forward(self, hidden_states, padded=False):
        """Logits [..., vocab_size]; with padded=True the kernel's own row layout [..., pad8(vocab_size)] whose extra
        columns are -inf — a softmax / cross entropy over it equals the one over vocab_size classes, and neither a
        compaction copy of the 30522-wide logits (forward) nor a re-padding copy of their gradient (backward) is
        needed."""
        h = self.transform(hidden_states)
        shp = h.shape
        logits = XF.linear(h.reshape(-1, shp[-1]), self._decoder_shadow(), self.bias, pad_value=float("-inf") if padded else None)
        return logits.reshape(*shp[:-1], -1)


    def loss_rows(self, hidden_states, labels):
        """Per-position cross entropy (0 where the label is negative / -100) of transform + tied decoder WITHOUT storing
        the [positions, vocab] logits: vocabulary GEMM fused with an online-softmax cross entropy (functional._VocabCEFn).
        hidden_states [..., D], labels [...] -> [positions] fp32."""
        h = self.transform(hidden_states)
        return XF.vocab_cross_entropy(h.reshape(-1, h.shape[-1]), self._decoder_shadow(), self.bias, labels.reshape(-1))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


class BertPreTrainedModel(nn.Module):
    """Weight init of the reference's BertPreTrainedModel (models/xbert.py:859-881) without the HF
    `PreTrainedModel` machinery (from_pretrained is not used by X2-VLM: weights arrive via load_state_dict)."""
    config_class = BertConfig
    base_model_prefix = "bert"

    def __init__(self, config):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)
        self.tie_weights()

    def tie_weights(self):
        out = self.get_output_embeddings() if hasattr(self, "get_output_embeddings") else None
        if out is not None and getattr(self.config, "tie_word_embeddings", True):
            out.weight = self.get_input_embeddings().weight

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def get_head_mask(self, head_mask, num_hidden_layers, *a, **k):
        if head_mask is not None:
            raise NotImplementedError("head_mask")
        return [None] * num_hidden_layers

    def get_extended_attention_mask(self, attention_mask, input_shape, device, is_decoder):
        """(1 - m) * -10000, broadcastable to [B, heads, Lq, Lk]; causal when is_decoder (models/xbert.py:1013-1073)."""
        if attention_mask.dim() == 3:
            ext = attention_mask[:, None, :, :]
        elif attention_mask.dim() == 2:
            if is_decoder:
                batch_size, seq_length = input_shape
                seq_ids = torch.arange(seq_length, device=device)
                causal = (seq_ids[None, None, :].repeat(batch_size, seq_length, 1) <= seq_ids[None, :, None]).to(attention_mask.dtype)
                if causal.shape[1] < attention_mask.shape[1]:
                    prefix = attention_mask.shape[1] - causal.shape[1]
                    causal = torch.cat([torch.ones((batch_size, seq_length, prefix), device=device, dtype=causal.dtype), causal], dim=-1)
                ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
            else:
                ext = attention_mask[:, None, None, :]
        else:
            raise ValueError("Wrong shape for input_ids (shape {}) or attention_mask (shape {})".format(
                input_shape, attention_mask.shape))
        return (1.0 - ext.to(dtype=torch.float32)) * -10000.0

    def invert_attention_mask(self, encoder_attention_mask):
        """Additive mask for cross-attention keys (HF invert_attention_mask; any large negative value gives
        identical probabilities as long as one key is visible — SURVEY.md §8c)."""
        if encoder_attention_mask.dim() == 3:
            ext = encoder_attention_mask[:, None, :, :]
        elif encoder_attention_mask.dim() == 2:
            ext = encoder_attention_mask[:, None, None, :]
        else:
            raise ValueError("encoder_attention_mask must be 2-D or 3-D")
        return (1.0 - ext.to(dtype=torch.float32)) * -10000.0


class BertModel(BertPreTrainedModel):
    def __init__(self, config, add_pooling_layer=True, add_embeddings_layer=True):
        super().__init__(config)
        if add_embeddings_layer:
            self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        self.init_weights()

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None,
                past_key_values=None, history_states=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, is_decoder=False, mode='multi_modal', split_lengths=None, encoder_kv_index=None):
        output_attentions = output_attentions if output_attentions is not None else getattr(self.config, "output_attentions", False)
        output_hidden_states = output_hidden_states if output_hidden_states is not None else getattr(self.config, "output_hidden_states", False)
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        if is_decoder:
            use_cache = use_cache if use_cache is not None else getattr(self.config, "use_cache", True)
            if use_cache and torch.is_grad_enabled():
                use_cache = False  # the cache is a forward-only product; training never reads it (xbert.py:1349-1350)
        else:
            use_cache = False
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        elif input_ids is not None:
            input_shape, device = input_ids.size(), input_ids.device
        elif inputs_embeds is not None:
            input_shape, device = inputs_embeds.size()[:-1], inputs_embeds.device
        elif encoder_embeds is not None:
            input_shape, device = encoder_embeds.size()[:-1], encoder_embeds.device
        else:
            raise ValueError("You have to specify either input_ids or inputs_embeds or encoder_embeds")
        batch_size, seq_length = input_shape
        past_key_values_length = past_key_values[0][0].shape[2] if past_key_values is not None else 0
        if attention_mask is None:
            attention_mask = torch.ones((batch_size, seq_length + past_key_values_length), device=device)
        if token_type_ids is None:
            token_type_ids = torch.zeros(input_shape, dtype=torch.long, device=device)
        extended_attention_mask = self.get_extended_attention_mask(attention_mask, input_shape, device, is_decoder)
        if encoder_hidden_states is not None:
            if isinstance(encoder_hidden_states, list):
                raise ValueError("list encoder_hidden_states are not supported")
            n_enc, enc_len, _ = encoder_hidden_states.size()
            if encoder_attention_mask is None:
                encoder_attention_mask = torch.ones((batch_size if encoder_kv_index is not None else n_enc, enc_len), device=device)
            encoder_extended_attention_mask = self.invert_attention_mask(encoder_attention_mask)
        else:
            encoder_extended_attention_mask = None
        head_mask = self.get_head_mask(head_mask, self.config.num_hidden_layers)
        if encoder_embeds is None:
            embedding_output = self.embeddings(input_ids=input_ids, position_ids=position_ids, token_type_ids=token_type_ids,
                                               inputs_embeds=inputs_embeds, past_key_values_length=past_key_values_length)
        else:
            embedding_output = encoder_embeds
        encoder_outputs = self.encoder(embedding_output, attention_mask=extended_attention_mask, head_mask=head_mask,
                                       encoder_hidden_states=encoder_hidden_states,
                                       encoder_attention_mask=encoder_extended_attention_mask,
                                       past_key_values=past_key_values, history_states=history_states, use_cache=use_cache,
                                       output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                                       return_dict=return_dict, mode=mode, split_lengths=split_lengths,
                                       encoder_kv_index=encoder_kv_index)
        sequence_output = encoder_outputs[0]
        pooled_output = self.pooler(sequence_output) if self.pooler is not None else None
        if not return_dict:
            return (sequence_output, pooled_output) + encoder_outputs[1:]
        return BaseModelOutputWithPoolingAndCrossAttentions(
            last_hidden_state=sequence_output, pooler_output=pooled_output, past_key_values=encoder_outputs.past_key_values,
            hidden_states=encoder_outputs.hidden_states, attentions=encoder_outputs.attentions,
            cross_attentions=encoder_outputs.cross_attentions)


class LabelSmoothSoftmaxCEV1(nn.Module):
    """Label-smoothed cross entropy with ignore_index (models/xbert.py:1223-1263): target distribution
    (1 - s) on the label and s / C on every class; ignored rows contribute 0 and are excluded from the mean."""

    def __init__(self, lb_smooth=0.1, reduction='mean', ignore_index=-100):
        super().__init__()
        self.lb_smooth, self.reduction, self.lb_ignore = lb_smooth, reduction, ignore_index

    def forward(self, logits, label):
        logs = F.log_softmax(logits.float(), dim=1)
        ignore = label.eq(self.lb_ignore)
        safe = label.masked_fill(ignore, 0)
        n_cls = logits.size(1)
        lb_pos, lb_neg = 1.0 - self.lb_smooth, self.lb_smooth / n_cls
        # -sum_c t_c * logp_c with t = lb_neg everywhere, lb_pos on the label (which overwrites, not adds to, lb_neg)
        loss = -(lb_neg * logs.sum(dim=1) + (lb_pos - lb_neg) * logs.gather(1, safe.unsqueeze(1)).squeeze(1))
        loss = loss.masked_fill(ignore, 0.0)
        if self.reduction == 'mean':
            return loss.sum() / ignore.eq(0).sum()
        if self.reduction == 'sum':
            return loss.sum()
        return loss


class BertLMHeadModel(BertPreTrainedModel):
    """BERT decoder with the LM head (models/xbert.py:1268-1413): VQA answer decoder and captioning language model
    (models/model_generation.py:443,647).  Causal masking comes from `is_decoder=True`; `labels` give the shifted
    next-token loss with optional label smoothing and `reduction='none'` (per-sequence sums)."""

    def __init__(self, config, label_smoothing=0.0):
        super().__init__(config)
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        self.label_smoothing = label_smoothing
        self.init_weights()
        self.bert.embeddings.word_embeddings.weight._x2k_autograd_too = True  # embedding lookup + tied decoder GEMM

    def get_input_embeddings(self):
        return self.bert.embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def set_output_embeddings(self, new_embeddings):
        self.cls.predictions.decoder = new_embeddings

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, labels=None,
                past_key_values=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                is_decoder=True, reduction='mean', mode='multi_modal', return_logits=False):
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        if labels is not None:
            use_cache = False
        outputs = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                            head_mask=head_mask, inputs_embeds=inputs_embeds, encoder_hidden_states=encoder_hidden_states,
                            encoder_attention_mask=encoder_attention_mask, past_key_values=past_key_values,
                            use_cache=use_cache, output_attentions=output_attentions,
                            output_hidden_states=output_hidden_states, return_dict=return_dict, is_decoder=is_decoder, mode=mode)
        sequence_output = outputs[0]
        if (labels is not None and not return_logits and self.label_smoothing <= 0 and getattr(self.config, "x2k_fused_ce", True)
                and sequence_output.shape[1] > 1):
            # shifted next-token loss (xbert.py:1359-1371) with the vocabulary GEMM fused into the cross entropy: position
            # t predicts token t+1, the last position predicts nothing; `logits` of the output is None on this path
            B_, L_ = sequence_output.shape[:2]
            rows = self.cls.predictions.loss_rows(sequence_output[:, :-1].contiguous(), labels[:, 1:].contiguous())
            if reduction == 'none':
                lm_loss = rows.view(B_, L_ - 1).sum(1)
            elif reduction == 'sum':
                lm_loss = rows.sum()
            else:
                lm_loss = rows.sum() / (labels[:, 1:] != -100).sum()
            if not return_dict:
                return (lm_loss, None) + tuple(outputs[2:])
            return CausalLMOutputWithCrossAttentions(loss=lm_loss, logits=None, past_key_values=outputs.past_key_values,
                                                     hidden_states=outputs.hidden_states, attentions=outputs.attentions,
                                                     cross_attentions=outputs.cross_attentions)
        prediction_scores = self.cls(sequence_output)
        if return_logits:
            return prediction_scores[:, :-1, :].contiguous()
        lm_loss = None
        if labels is not None:
            shifted = prediction_scores[:, :-1, :].contiguous()
            labels = labels[:, 1:].contiguous()
            if self.label_smoothing > 0:
                loss_fct = LabelSmoothSoftmaxCEV1(lb_smooth=self.label_smoothing, reduction=reduction)
            else:
                loss_fct = nn.CrossEntropyLoss(reduction=reduction)
            lm_loss = loss_fct(shifted.view(-1, self.config.vocab_size), labels.view(-1))
            if reduction == 'none':
                lm_loss = lm_loss.view(prediction_scores.size(0), -1).sum(1)
        if not return_dict:
            output = (prediction_scores,) + outputs[2:]
            return ((lm_loss,) + output) if lm_loss is not None else output
        return CausalLMOutputWithCrossAttentions(loss=lm_loss, logits=prediction_scores, past_key_values=outputs.past_key_values,
                                                 hidden_states=outputs.hidden_states, attentions=outputs.attentions,
                                                 cross_attentions=outputs.cross_attentions)

    def prepare_inputs_for_generation(self, input_ids, past=None, attention_mask=None, **model_kwargs):
        """Cut the prompt to its last token once a cache exists (models/xbert.py:1390-1407)."""
        if attention_mask is None:
            attention_mask = input_ids.new_ones(input_ids.shape)
        if past is not None:
            input_ids = input_ids[:, -1:]
        return {"input_ids": input_ids, "attention_mask": attention_mask, "past_key_values": past,
                "encoder_hidden_states": model_kwargs.get("encoder_hidden_states", None),
                "encoder_attention_mask": model_kwargs.get("encoder_attention_mask", None), "is_decoder": True}

    def _reorder_cache(self, past, beam_idx):
        return tuple(tuple(t.index_select(0, beam_idx) for t in layer_past) for layer_past in past)

    @torch.no_grad()
    def generate_no_beam(self, input_ids, max_length, do_sample=False, temperature=1.0, top_k=0, top_p=1.0,
                         repetition_penalty=1.0, pad_token_id=0, eos_token_ids=(), **model_kwargs):
        """The num_beams == 1 generation loop of the reference (models/xbert.py:1415-1519) on the key/value cache:
        CTRL repetition penalty, temperature, top-k / top-p sampling or arg-max, finished sequences padded, the last
        position forced to EOS when max_length is hit.  Returns (ids [B, max_length], mean log-prob per sequence)."""
        batch_size = input_ids.shape[0]
        cur_unfinished = input_ids.new_ones(batch_size)
        logprobs, unfinished_sents = [], []
        past, cur_len = None, input_ids.shape[1]
        while cur_len < max_length:
            inp = self.prepare_inputs_for_generation(input_ids, past=past, **model_kwargs)
            out = self(**inp, use_cache=True, return_dict=True)
            past = out.past_key_values
            next_token_logits = out.logits[:, -1, :].float().clone()
            if repetition_penalty != 1.0:
                seen = torch.zeros_like(next_token_logits, dtype=torch.bool).scatter_(1, input_ids, True)
                scaled = torch.where(next_token_logits < 0, next_token_logits * repetition_penalty,
                                     next_token_logits / repetition_penalty)
                next_token_logits = torch.where(seen, scaled, next_token_logits)
            if do_sample:
                if temperature != 1.0:
                    next_token_logits = next_token_logits / temperature
                next_token_logits = top_k_top_p_filtering(next_token_logits, top_k=top_k, top_p=top_p)
                next_token = torch.multinomial(F.softmax(next_token_logits, dim=-1), num_samples=1).squeeze(1)
            else:
                next_token = torch.argmax(next_token_logits, dim=-1)
            logprobs.append(torch.gather(F.log_softmax(next_token_logits, dim=-1), -1, next_token.unsqueeze(-1)))
            unfinished_sents.append(cur_unfinished)
            tokens_to_add = next_token * cur_unfinished + pad_token_id * (1 - cur_unfinished)
            input_ids = torch.cat([input_ids, tokens_to_add.unsqueeze(-1)], dim=-1)
            if "attention_mask" in model_kwargs and model_kwargs["attention_mask"] is not None:
                am = model_kwargs["attention_mask"]
                model_kwargs["attention_mask"] = torch.cat([am, am.new_ones((am.shape[0], 1))], dim=-1)
            cur_len += 1
            for eos in eos_token_ids:
                cur_unfinished = cur_unfinished.mul(tokens_to_add.ne(eos).long())
            if cur_unfinished.max() == 0:
                break
        if cur_len == max_length and len(eos_token_ids) > 0:
            input_ids[:, -1].masked_fill_(cur_unfinished.to(dtype=torch.bool), eos_token_ids[0])
        logprobs = torch.cat(logprobs, dim=1)
        unfinished = torch.stack(unfinished_sents, dim=1).float()
        mean_logprob = (logprobs * unfinished).sum(dim=1) / unfinished.sum(dim=1)
        if max_length > input_ids.shape[1]:
            pad = input_ids.new_full((batch_size, max_length - input_ids.shape[1]), pad_token_id)
            input_ids = torch.cat([input_ids, pad], dim=1)
        return input_ids, mean_logprob

    @torch.no_grad()
    def greedy_decode(self, input_ids, max_length, eos_token_id=None, pad_token_id=0, **model_kwargs):
        """Cached greedy decoding (the num_beams == 1, do_sample == False case of the reference's generate loop,
        models/xbert.py:1415-1490): one prompt pass that fills the cache, then one token per step."""
        past, cur = None, input_ids
        unfinished = torch.ones(input_ids.shape[0], dtype=torch.long, device=input_ids.device)
        while cur.shape[1] < max_length:
            inp = self.prepare_inputs_for_generation(cur, past=past, **model_kwargs)
            out = self(**inp, use_cache=True, return_dict=True)
            past = out.past_key_values
            nxt = out.logits[:, -1, :].argmax(dim=-1)
            if eos_token_id is not None:
                nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
                unfinished = unfinished * nxt.ne(eos_token_id).long()
            cur = torch.cat([cur, nxt[:, None]], dim=1)
            if eos_token_id is not None and unfinished.max() == 0:
                break
        return cur


def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("Inf"), min_tokens_to_keep=1):
    """Top-k and / or nucleus filtering of a [batch, vocab] logits matrix, in place (models/xbert.py:1521-1554): keep the
    top_k highest logits; of the rest keep the smallest prefix (by descending probability) whose cumulative probability
    reaches top_p, always including the token that crosses the threshold and at least min_tokens_to_keep tokens."""
    if top_k > 0:
        k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits.masked_fill_(logits < kth, filter_value)
    if top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        drop_sorted = cum > top_p
        if min_tokens_to_keep > 1:
            drop_sorted[..., :min_tokens_to_keep] = False
        drop_sorted = torch.cat([torch.zeros_like(drop_sorted[..., :1]), drop_sorted[..., :-1]], dim=-1)  # shift right
        drop = torch.zeros_like(drop_sorted).scatter(1, sorted_idx, drop_sorted)
        logits.masked_fill_(drop, filter_value)
    return logits


class BertForMaskedLM(BertPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        self.init_weights()
        self.bert.embeddings.word_embeddings.weight._x2k_autograd_too = True  # embedding lookup + tied decoder GEMM

    def get_input_embeddings(self):
        return self.bert.embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def set_output_embeddings(self, new_embeddings):
        self.cls.predictions.decoder = new_embeddings

    def gather_seq_out_by_pos(self, seq, pos):
        return torch.gather(seq, 1, pos.unsqueeze(2).expand(-1, -1, seq.size(-1)))

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, labels=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, is_decoder=False, mode='multi_modal',
                return_logits=False, masked_pos=None, reduction='mean', split_lengths=None, past_key_values=None,
                encoder_kv_index=None):
        return_dict = return_dict if return_dict is not None else getattr(self.config, "use_return_dict", True)
        outputs = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                            head_mask=head_mask, inputs_embeds=inputs_embeds, encoder_embeds=encoder_embeds,
                            encoder_hidden_states=encoder_hidden_states, encoder_attention_mask=encoder_attention_mask,
                            output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                            return_dict=return_dict, is_decoder=is_decoder, mode=mode, split_lengths=split_lengths,
                            past_key_values=past_key_values, encoder_kv_index=encoder_kv_index)
        sequence_output = outputs[0]
        if masked_pos is None:
            raise NotImplementedError("need check!")  # as in the reference (xbert.py:1650)
        sequence_output = self.gather_seq_out_by_pos(sequence_output, masked_pos)
        if labels is not None and not return_logits and getattr(self.config, "x2k_fused_ce", True):
            # loss only (what XVLMBase.get_mlm_loss reads, models/xvlm.py:900-908): the [B, n, 30522] logits are never
            # materialised — `logits` of the returned output is None; set config.x2k_fused_ce = False to get them back
            rows = self.cls.predictions.loss_rows(sequence_output, labels)
            masked_lm_loss = rows.sum() / (labels.reshape(-1) != -100).sum()
            if not return_dict:
                return (masked_lm_loss, None) + tuple(outputs[2:])
            return MaskedLMOutput(loss=masked_lm_loss, logits=None, hidden_states=outputs.hidden_states,
                                  attentions=outputs.attentions)
        padded_scores = self.cls.predictions(sequence_output, padded=True)       # [B, n, pad8(V)], -inf beyond V
        prediction_scores = padded_scores[..., :self.config.vocab_size]
        if return_logits:
            return prediction_scores
        masked_lm_loss = None
        if labels is not None:
            masked_lm_loss = F.cross_entropy(padded_scores.view(-1, padded_scores.shape[-1]), labels.view(-1))
        if not return_dict:
            output = (prediction_scores,) + outputs[2:]
            return ((masked_lm_loss,) + output) if masked_lm_loss is not None else output
        return MaskedLMOutput(loss=masked_lm_loss, logits=prediction_scores, hidden_states=outputs.hidden_states,
                              attentions=outputs.attentions)
