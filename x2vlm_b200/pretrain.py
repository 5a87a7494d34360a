"""X²-VLM pre-training model on the B200-native encoders.

`XVLM` mirrors the reference's models/model_pretrain.py:XVLM / models/xvlm.py:XVLMBase surface for the
pre-training path — same sub-module names (`vision_encoder`, `text_encoder`, `vision_proj`, `text_proj`,
`temp`, `itm_head`, `bbox_head`), same state_dict keys (587 for base), same method names and `forward`
keyword arguments — so reference checkpoints load and the reference's callers read the same losses.
(In the build container the reference's own XVLMBase is also constructed on top of these encoders,
tests/test_dropin_reference.py; the reference tree does not exist on the GPU box, hence this mirror.)

Differences that are part of the B200-first design, none of which change the training semantics:
  * hard negatives are drawn on the device (`torch.multinomial` over all rows at once) instead of
    2·B host-synchronising `.item()` calls (models/xvlm.py:847-855);
  * `forward_mixed` runs one image iteration + one region iteration (Pretrain.py:run_mixed_iter) with
    every pass through a stack batched into ONE call: 1 vision call (images of both sub-batches),
    1 text call (clean + masked captions of both), 1 fusion call (ITM pos / ITM neg / MLM / bbox of
    both) in which sequences that look at the same image share its K/V projection (`encoder_kv_index`).
"""

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F
from transformers import BertConfig

from . import beit2, xbert

BERT_BASE = dict(attention_probs_dropout_prob=0.1, hidden_act="gelu", hidden_dropout_prob=0.1, hidden_size=768,
                 initializer_range=0.02, intermediate_size=3072, layer_norm_eps=1e-12, max_position_embeddings=512,
                 num_attention_heads=12, num_hidden_layers=12, pad_token_id=0, type_vocab_size=2, vocab_size=30522)
BERT_LARGE = dict(BERT_BASE, hidden_size=1024, intermediate_size=4096, num_attention_heads=16)


def base_config(**over):
    """Shape keys of configs/pretrain/x2vlm_base_*.yaml."""
    cfg = dict(use_beit_v2=True, vision_config="configs/config_beit2_base.json", image_res=224, patch_size=16,
               text_encoder="data/bert-base-uncased", text_num_hidden_layers=18, text_fusion_start_at=12, embed_dim=256,
               temp=0.07)
    cfg.update(over)
    return cfg


def large_config(**over):
    """configs/pretrain/x2vlm_large_*.yaml: beit2-large + 12 text / 6 fusion layers of width 1024."""
    cfg = base_config(vision_config="configs/config_beit2_large.json", text_encoder="data/bert-large-uncased-12l")
    cfg.update(over)
    return cfg


class AllGather(torch.autograd.Function):
    """all_gather whose backward keeps only the local slice of the gradient (models/xvlm.py:140-160)."""

    @staticmethod
    def forward(ctx, tensor, rank, world_size):
        output = [torch.empty_like(tensor) for _ in range(world_size)]
        dist.all_gather(output, tensor.contiguous())
        ctx.rank, ctx.batch_size = rank, tensor.shape[0]
        return torch.cat(output, 0)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output[ctx.batch_size * ctx.rank: ctx.batch_size * (ctx.rank + 1)], None, None


def allgather(t):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return AllGather.apply(t, dist.get_rank(), dist.get_world_size())
    return t


def allgather_packed(feats):
    """ONE all-gather for several [B_k, D] feature matrices of a step (the ITC features of the image and the region
    sub-batch): returns the gathered [W * B_k, D] matrices, rank-major exactly as one `allgather` per matrix gives them
    (models/xvlm.py:804-805); the backward keeps the local slices (`AllGather`)."""
    W = dist.get_world_size()
    packed = torch.cat(feats, dim=0)
    allp = allgather(packed).view(W, packed.shape[0], -1)
    out, o = [], 0
    for f in feats:
        n = f.shape[0]
        out.append(allp[:, o:o + n].reshape(W * n, -1))
        o += n
    return out


def build_mlp(input_dim, output_dim):
    return nn.Sequential(nn.Linear(input_dim, input_dim * 2), nn.LayerNorm(input_dim * 2), nn.GELU(),
                         nn.Linear(input_dim * 2, output_dim))


def box_cxcywh_to_xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def giou_pairs(b1, b2):
    """Row-wise generalized IoU of xyxy boxes (the diagonal of models/box_ops.py:generalized_box_iou)."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = (torch.min(b1[:, 2:], b2[:, 2:]) - torch.max(b1[:, :2], b2[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = a1 + a2 - inter
    whc = (torch.max(b1[:, 2:], b2[:, 2:]) - torch.min(b1[:, :2], b2[:, :2])).clamp(min=0)
    area = whc[:, 0] * whc[:, 1]
    return inter / union - (area - union) / area


class _GatherRowsFn(torch.autograd.Function):
    """rows = x.view(-1, D)[idx]; backward = one zero fill + one index_add (idx may repeat: padded mask slots)."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        return x.reshape(-1, x.shape[-1]).index_select(0, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        d = torch.zeros(ctx.shape, dtype=g.dtype, device=g.device)
        d.view(-1, ctx.shape[-1]).index_add_(0, idx, g)
        return d, None


def _gather_rows(x, idx):
    return _GatherRowsFn.apply(x, idx)


def _itm_labels(bs, device):
    """[1]*bs + [0]*2bs (models/xvlm.py:895-897), built on the device: a pageable host-to-device copy would be a
    hidden sync and is illegal inside CUDA-graph capture."""
    labels = torch.zeros(3 * bs, dtype=torch.long, device=device)
    labels[:bs] = 1
    return labels


class XVLM(nn.Module):
    def __init__(self, config, load_vision_params=False, load_text_params=False, pretraining=True):
        super().__init__()
        if load_vision_params or load_text_params:
            raise NotImplementedError("checkpoint loading goes through load_state_dict (no pretrained files offline)")
        large = "large" in config.get("vision_config", "")
        factory = beit2.beit_large_patch16 if large else beit2.beit_base_patch16
        self.vision_encoder = factory(img_size=config["image_res"], drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.0,
                                      use_mean_pooling=True, init_scale=0.001, use_rel_pos_bias=True, use_abs_pos_emb=False,
                                      init_values=0.1, qkv_bias=True, local_attn_depth=config.get("local_attn_depth", -1),
                                      vision_num_hidden_layers=config.get("vision_num_hidden_layers", -1))
        self.vision_width = self.vision_encoder.vision_width = self.vision_encoder.embed_dim
        tcfg = BertConfig(**(BERT_LARGE if "large" in config.get("text_encoder", "") else BERT_BASE))
        tcfg.num_hidden_layers = config["text_num_hidden_layers"]
        tcfg.fusion_layer = config["text_fusion_start_at"]
        tcfg.embedding_dim = tcfg.hidden_size
        tcfg.hidden_dropout_prob = config.get("dropout", tcfg.hidden_dropout_prob)
        tcfg.encoder_width = self.vision_width
        tcfg.text_drop_path_rate = config.get("text_drop_path_rate", 0.0)
        tcfg.cross_drop_path_rate = config.get("cross_drop_path_rate", 0.0)
        self.text_encoder = xbert.BertForMaskedLM(config=tcfg)
        self.text_width = tcfg.hidden_size
        self.embed_dim = config["embed_dim"]
        self.vision_proj = nn.Linear(self.vision_width, self.embed_dim)
        self.text_proj = nn.Linear(self.text_width, self.embed_dim)
        self.temp = nn.Parameter(torch.ones([]) * config["temp"])
        self.itm_head = build_mlp(input_dim=self.text_width, output_dim=2)
        self.bbox_head = build_mlp(input_dim=self.text_width, output_dim=4)
        self.init_params = []
        # video input (models/xvlm.py:483-501): frames go through the image encoder, then a mean over time
        self.video_encoding = config.get("video_encoding", "")
        if self.video_encoding not in ("", "avgpool"):
            raise NotImplementedError("video_encoding == %r (the reference raises for everything but 'avgpool' too, "
                                      "models/xvlm.py:651-654)" % (self.video_encoding,))
        if self.video_encoding:
            self.frame_len = config["frame_len"]
            self.add_frame_pos = config["add_frame_pos"]
            if self.add_frame_pos:
                self.absolute_frame_pos_embed = nn.Parameter(torch.zeros(1, self.frame_len, 1, self.vision_width))
                beit2._trunc_normal_(self.absolute_frame_pos_embed, std=.02)
                self.init_params.append("absolute_frame_pos_embed")

    # ------------------------------------------------------------------ reference-shaped methods
    def get_frame_embeds(self, frames):
        """models/xvlm.py:615-661, video_encoding == 'avgpool': every frame through the vision encoder as one batch of
        bsz * frame_len images, learned per-frame offset, mean over the frame axis.  frames: [bsz, F, 3, H, W]."""
        assert frames.dim() == 5 and self.video_encoding == "avgpool"
        bsz, n_f = frames.shape[:2]
        emb = self.vision_encoder(frames.reshape(bsz * n_f, *frames.shape[2:]))
        emb = emb.view(bsz, n_f, emb.shape[1], emb.shape[2])
        if self.add_frame_pos:
            emb = emb + self.absolute_frame_pos_embed
        emb = emb.mean(dim=1)
        return emb, torch.ones(emb.size()[:-1], dtype=torch.long, device=frames.device)

    def get_vision_embeds(self, image, image_atts=None, idx_to_group_img=None):
        """models/xvlm.py:663-713."""
        if image.dim() == 5:
            assert idx_to_group_img is None, "not supported"
            return self.get_frame_embeds(image)
        if idx_to_group_img is None:
            image_embeds = self.vision_encoder(image)
            return image_embeds, torch.ones(image_embeds.size()[:-1], dtype=torch.long, device=image.device)
        if image_atts is None:  # fewer images than samples, full attention (models/xvlm.py:679-690)
            full = self.vision_encoder(image)[idx_to_group_img]
            return full, torch.ones(full.size()[:-1], dtype=torch.long, device=image.device)
        image_embeds, full = self.vision_encoder(image, idx_to_group_img=idx_to_group_img, image_atts=image_atts)
        return image_embeds, image_atts, full[idx_to_group_img]

    def get_text_embeds(self, text_ids, text_atts):
        return self.text_encoder.bert(text_ids, attention_mask=text_atts, return_dict=True, mode='text').last_hidden_state

    def get_cross_embeds(self, image_embeds, image_atts, text_ids=None, text_embeds=None, text_atts=None,
                         encoder_kv_index=None):
        assert text_atts is not None
        enc = self.text_encoder.bert
        if text_embeds is not None:
            return enc(encoder_embeds=text_embeds, attention_mask=text_atts, encoder_hidden_states=image_embeds,
                       encoder_attention_mask=image_atts, return_dict=True, mode='fusion',
                       encoder_kv_index=encoder_kv_index).last_hidden_state
        if text_ids is not None:
            return enc(text_ids, attention_mask=text_atts, encoder_hidden_states=image_embeds,
                       encoder_attention_mask=image_atts, return_dict=True,
                       encoder_kv_index=encoder_kv_index).last_hidden_state
        raise ValueError

    def get_features(self, image_embeds=None, text_embeds=None):
        if image_embeds is None:
            return F.normalize(self.text_proj(text_embeds[:, 0, :]), dim=-1)
        if text_embeds is None:
            return F.normalize(self.vision_proj(image_embeds[:, 0, :]), dim=-1)
        return F.normalize(self.vision_proj(image_embeds[:, 0, :]), dim=-1), \
            F.normalize(self.text_proj(text_embeds[:, 0, :]), dim=-1)

    def get_contrastive_loss(self, image_feat, text_feat, idx=None, gathered=None):
        """models/xvlm.py:794-826.  idx [B] (retrieval fine-tuning): samples with the same id are positives of each
        other — soft labels = row-normalised id-equality matrix over the gathered batch.  `gathered` = (image_feat_all,
        text_feat_all) when the caller has already all-gathered the features (one collective for several losses)."""
        image_feat_all, text_feat_all = gathered if gathered is not None else (allgather(image_feat), allgather(text_feat))
        logits = image_feat_all @ text_feat_all.t() / self.temp
        if idx is None:
            labels = torch.arange(logits.shape[0], device=logits.device)
            return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2
        idx = idx.view(-1, 1)
        assert idx.size(0) == image_feat.size(0)
        idx_all = allgather(idx)
        pos_idx = torch.eq(idx_all, idx_all.t()).float()
        labels = pos_idx / pos_idx.sum(1, keepdim=True)
        loss_i2t = -torch.sum(F.log_softmax(logits, dim=1) * labels, dim=1).mean()
        loss_t2i = -torch.sum(F.log_softmax(logits.t(), dim=1) * labels, dim=1).mean()
        return (loss_i2t + loss_t2i) / 2

    def hard_negative_weights(self, image_feat, text_feat, idx=None):
        """Sampling weights of models/xvlm.py:828-846: softmax(sim) + 1e-5 with the positives zeroed — the diagonal, or
        every pair that shares an id when idx is given.  Returns (weights_i2t, weights_t2i)."""
        with torch.no_grad():
            sim_i2t = image_feat @ text_feat.t() / self.temp
            w_i2t = F.softmax(sim_i2t, dim=1) + 1e-5
            w_t2i = F.softmax(sim_i2t.t(), dim=1) + 1e-5
            if idx is None:
                w_i2t.fill_diagonal_(0)
                w_t2i.fill_diagonal_(0)
            else:
                idx = idx.view(-1, 1)
                assert idx.size(0) == image_feat.size(0)
                mask = torch.eq(idx, idx.t())
                w_i2t.masked_fill_(mask, 0)
                w_t2i.masked_fill_(mask, 0)
        return w_i2t, w_t2i

    def get_hard_negatives(self, image_feat, text_feat, idx=None):
        """Same sampling law as models/xvlm.py:828-857, drawn for all rows in one device call; returns index TENSORS
        (no host sync)."""
        w_i2t, w_t2i = self.hard_negative_weights(image_feat, text_feat, idx)
        with torch.no_grad():
            image_neg_idx = torch.multinomial(w_t2i, 1).squeeze(1)
            text_neg_idx = torch.multinomial(w_i2t, 1).squeeze(1)
        return image_neg_idx, text_neg_idx

    def get_matching_loss(self, image_embeds, image_atts, image_feat, text_embeds, text_atts, text_feat, idx=None,
                          neg_idx=None):
        """models/xvlm.py:859-899; the three ITM passes run as one fusion batch of 3·bs sequences that index the bs
        image K/V sets."""
        image_neg_idx, text_neg_idx = neg_idx if neg_idx is not None else self.get_hard_negatives(image_feat, text_feat, idx)
        bs = image_feat.size(0)
        ar = torch.arange(bs, device=image_feat.device)
        kv_index = torch.cat([ar, image_neg_idx, ar]).to(torch.int32)
        text_all = torch.cat([text_embeds, text_embeds, text_embeds[text_neg_idx]], dim=0)
        tatt_all = torch.cat([text_atts, text_atts, text_atts[text_neg_idx]], dim=0)
        iatt_all = torch.cat([image_atts, image_atts[image_neg_idx], image_atts], dim=0)
        cross = self.get_cross_embeds(image_embeds, iatt_all, text_embeds=text_all, text_atts=tatt_all,
                                      encoder_kv_index=kv_index)[:, 0, :]
        output = self.itm_head(cross)
        itm_labels = _itm_labels(bs, output.device)
        return F.cross_entropy(output, itm_labels)

    def get_mlm_loss(self, text_ids_masked, text_atts, image_embeds, image_atts, masked_pos, masked_ids):
        return self.text_encoder(text_ids_masked, attention_mask=text_atts, encoder_hidden_states=image_embeds,
                                 encoder_attention_mask=image_atts, return_dict=True, labels=masked_ids,
                                 masked_pos=masked_pos).loss

    def predict_bbox(self, image_embeds, text_embeds, text_atts):
        assert image_embeds.size(0) == text_embeds.size(0)
        cls = self.get_cross_embeds(image_embeds, torch.ones(image_embeds.shape[:2], device=image_embeds.device),
                                    text_embeds=text_embeds, text_atts=text_atts)[:, 0, :]
        return self.bbox_head(cls).sigmoid()

    def get_bbox_loss(self, output_coord, target_bbox, is_image=None):
        """L1 + GIoU; degenerate boxes zero the GIoU term without a host sync (models/xvlm.py:927-957)."""
        loss_bbox = F.l1_loss(output_coord, target_bbox, reduction='none')
        b1, b2 = box_cxcywh_to_xyxy(output_coord), box_cxcywh_to_xyxy(target_bbox)
        degenerate = ((b1[:, 2:] < b1[:, :2]).any() | (b2[:, 2:] < b2[:, :2]).any())
        loss_giou = torch.where(degenerate, torch.zeros_like(b1[:, 0]), 1 - giou_pairs(b1, b2))
        if is_image is None:
            num_boxes = target_bbox.size(0)
        else:
            num_boxes = torch.sum(1 - is_image)
            loss_bbox = loss_bbox * (1 - is_image.view(-1, 1))
            loss_giou = loss_giou * (1 - is_image)
        return loss_bbox.sum() / num_boxes, loss_giou.sum() / num_boxes

    def forward(self, image=None, text_ids=None, text_atts=None, text_ids_masked=None, masked_pos=None, masked_ids=None,
                image_atts=None, idx_to_group_img=None, target_bbox=None, is_image=None, ret_bbox_loss=False,
                ret_match_loss=True, neg_idx=None):
        """models/model_pretrain.py:30-88 (image given), pass by pass like the reference."""
        if image is None:
            return {'loss_mlm': self.get_mlm_loss(text_ids_masked, text_atts, None, None, masked_pos, masked_ids)}
        if ret_bbox_loss:
            image_embeds, image_atts, image_embeds_fullatts = self.get_vision_embeds(image, image_atts, idx_to_group_img)
        else:
            image_embeds, image_atts = self.get_vision_embeds(image)
        text_embeds = self.get_text_embeds(text_ids, text_atts)
        image_feat, text_feat = self.get_features(image_embeds, text_embeds)
        loss = {'loss_itc': self.get_contrastive_loss(image_feat, text_feat)}
        if ret_match_loss:
            loss['loss_itm'] = self.get_matching_loss(image_embeds, image_atts, image_feat, text_embeds, text_atts, text_feat,
                                                      neg_idx=neg_idx)
        else:
            loss['loss_itm'] = torch.tensor(0.0)
        loss['loss_mlm'] = self.get_mlm_loss(text_ids_masked, text_atts, image_embeds, image_atts, masked_pos, masked_ids)
        if ret_bbox_loss:
            coord = self.predict_bbox(image_embeds_fullatts, text_embeds, text_atts)
            loss['loss_bbox'], loss['loss_giou'] = self.get_bbox_loss(coord, target_bbox, is_image=is_image)
        return loss

    def forward_retrieval(self, image, text_ids, text_atts, idx=None):
        """Retrieval fine-tuning step (models/model_retrieval.py:13-24): ITC + ITM, sample ids `idx` mark captions of
        the same image as mutual positives.  Returns (loss_itc, loss_itm)."""
        image_embeds, image_atts = self.get_vision_embeds(image)
        text_embeds = self.get_text_embeds(text_ids, text_atts)
        with torch.no_grad():
            self.temp.clamp_(0.001, 0.5)
        image_feat, text_feat = self.get_features(image_embeds, text_embeds)
        loss_itc = self.get_contrastive_loss(image_feat, text_feat, idx=idx)
        loss_itm = self.get_matching_loss(image_embeds, image_atts, image_feat, text_embeds, text_atts, text_feat, idx=idx)
        return loss_itc, loss_itm

    # ------------------------------------------------------------------ batched mixed step
    def forward_mixed(self, ib, rb=None, neg_idx_i=None, neg_idx_r=None, out=None):
        """One image iteration (+ one region iteration) of Pretrain.py:run_mixed_iter with every stack called once.
        ib / rb: dicts as produced by x2vlm_b200.synth.{image_text_batch, region_batch} (tensors on the device).
        Returns {'image': losses, 'region': losses}."""
        dev = ib["image"].device
        bert = self.text_encoder.bert
        Bi = ib["image"].shape[0]
        has_r = rb is not None
        Br = rb["text_ids"].shape[0] if has_r else 0
        n_img_r = rb["image"].shape[0] if has_r else 0
        # ---- vision: all images once ----
        if ib["image"].dim() == 5:  # video iteration (Pretrain.py:run_video_iter): frames -> per-frame encoding -> avgpool
            assert not has_r, "video batches carry no region sub-batch"
            full, _ = self.get_frame_embeds(ib["image"])
        else:
            images = torch.cat([ib["image"], rb["image"]]) if has_r else ib["image"]
            blocks_out, _ = self.vision_encoder.forward_blocks(images)
            full = self.vision_encoder.tail(blocks_out)  # [Bi + n_img_r, N, D]
        emb_i = full[:Bi]
        N = full.shape[1]
        atts_i = torch.ones(Bi, N, dtype=torch.long, device=dev)
        if has_r:
            emb_r = self.vision_encoder.tail(blocks_out[Bi:], rb["idx_to_group_img"], rb["image_atts"])
            atts_r = rb["image_atts"]
        # ---- text layers: clean + masked captions of both sub-batches once ----
        ids = [ib["text_ids"], ib["text_ids_masked"]] + ([rb["text_ids"], rb["text_ids_masked"]] if has_r else [])
        tat = [ib["text_atts"], ib["text_atts"]] + ([rb["text_atts"], rb["text_atts"]] if has_r else [])
        tatts_all = torch.cat(tat)
        t_all = bert(torch.cat(ids), attention_mask=tatts_all, return_dict=True, mode='text').last_hidden_state
        te_i, tm_i = t_all[:Bi], t_all[Bi:2 * Bi]
        if has_r:
            te_r, tm_r = t_all[2 * Bi:2 * Bi + Br], t_all[2 * Bi + Br:]
        # ---- ITC + hard negatives ----
        losses = {"image": {}, "region": {}}
        fi_i, ft_i = self.get_features(emb_i, te_i)
        if has_r:
            fi_r, ft_r = self.get_features(emb_r, te_r)
        # ONE all-gather for the (up to) four ITC feature matrices of the step instead of one per matrix
        g_i = g_r = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            g = allgather_packed([fi_i, ft_i] + ([fi_r, ft_r] if has_r else []))
            g_i = (g[0], g[1])
            if has_r:
                g_r = (g[2], g[3])
        losses["image"]["loss_itc"] = self.get_contrastive_loss(fi_i, ft_i, gathered=g_i)
        ineg_i, tneg_i = neg_idx_i if neg_idx_i is not None else self.get_hard_negatives(fi_i, ft_i)
        if has_r:
            losses["region"]["loss_itc"] = self.get_contrastive_loss(fi_r, ft_r, gathered=g_r)
            ineg_r, tneg_r = neg_idx_r if neg_idx_r is not None else self.get_hard_negatives(fi_r, ft_r)
        # ---- one fusion call: [ITM pos | ITM img-neg | ITM txt-neg | MLM] (image) + same (region) + bbox ----
        ar_i = torch.arange(Bi, device=dev)
        txt = [te_i, te_i, te_i[tneg_i], tm_i]
        tatt = [ib["text_atts"], ib["text_atts"], ib["text_atts"][tneg_i], ib["text_atts"]]
        kv = [ar_i, ineg_i, ar_i, ar_i]
        iatt = [atts_i, atts_i[ineg_i], atts_i, atts_i]
        enc = [emb_i]
        if has_r:
            ar_r = torch.arange(Br, device=dev)
            txt += [te_r, te_r, te_r[tneg_r], tm_r, te_r]
            tatt += [rb["text_atts"], rb["text_atts"], rb["text_atts"][tneg_r], rb["text_atts"], rb["text_atts"]]
            kv += [Bi + ar_r, Bi + ineg_r, Bi + ar_r, Bi + ar_r, Bi + Br + rb["idx_to_group_img"]]
            iatt += [atts_r, atts_r[ineg_r], atts_r, atts_r, torch.ones(Br, N, dtype=torch.long, device=dev)]
            enc += [emb_r, full[Bi:]]
        cross = self.get_cross_embeds(torch.cat(enc), torch.cat(iatt), text_embeds=torch.cat(txt), text_atts=torch.cat(tatt),
                                      encoder_kv_index=torch.cat(kv).to(torch.int32))
        # Every head reads a few rows of the [576, L, D] fusion output (cls tokens, masked positions).  Slicing it five
        # times makes autograd build five full-size zero gradients and add them up (9 passes over 70 MB); ONE gather of
        # all the rows the heads need has a backward of one fill + one index_add.
        L = cross.shape[1]
        o = 4 * Bi
        mpos, mids = [ib["masked_pos"]], [ib["masked_ids"]]
        cls_seq = [torch.arange(0, 3 * Bi, device=dev)]
        mlm_rows = [((3 * Bi + ar_i) * L).unsqueeze(1) + ib["masked_pos"]]
        if has_r:
            cls_seq += [torch.arange(o, o + 3 * Br, device=dev), torch.arange(o + 4 * Br, o + 5 * Br, device=dev)]
            mlm_rows.append(((o + 3 * Br + ar_r) * L).unsqueeze(1) + rb["masked_pos"])
            mpos.append(rb["masked_pos"]); mids.append(rb["masked_ids"])
        n_cls = 3 * Bi + (4 * Br if has_r else 0)
        rows = _gather_rows(cross, torch.cat([torch.cat(cls_seq) * L] + [m.reshape(-1) for m in mlm_rows]))
        cls_rows = rows[:n_cls]
        labels_i = _itm_labels(Bi, dev)
        itm_logits_i = self.itm_head(cls_rows[:3 * Bi])
        losses["image"]["loss_itm"] = F.cross_entropy(itm_logits_i, labels_i)
        if has_r:
            labels_r = _itm_labels(Br, dev)
            losses["region"]["loss_itm"] = F.cross_entropy(self.itm_head(cls_rows[3 * Bi:3 * Bi + 3 * Br]), labels_r)
            coord = self.bbox_head(cls_rows[3 * Bi + 3 * Br:]).sigmoid()
            losses["region"]["loss_bbox"], losses["region"]["loss_giou"] = self.get_bbox_loss(
                coord, rb["target_bbox"], is_image=rb["is_image"])
        # ---- MLM head once over the masked positions of both sub-batches ----
        n_pos = mpos[0].shape[1]
        seq = rows[n_cls:].view(-1, n_pos, rows.shape[-1])
        # vocabulary GEMM fused with the cross entropy: per-position losses (0 where the target is -100) without the
        # [positions, 30522] logits; then the two means
        tgt = torch.cat(mids).reshape(-1)
        ce = self.text_encoder.cls.predictions.loss_rows(seq, tgt)
        n_i = mids[0].numel()
        losses["image"]["loss_mlm"] = ce[:n_i].sum() / (tgt[:n_i] != -100).sum()
        if has_r:
            losses["region"]["loss_mlm"] = ce[n_i:].sum() / (tgt[n_i:] != -100).sum()
        if out is not None:
            with torch.no_grad():  # only on request (tests): the step itself never forms the logits
                logits = self.text_encoder.cls.predictions(seq[:Bi])
            out.update(image_embeds=emb_i, text_embeds=te_i, image_feat=fi_i, text_feat=ft_i, itm_logits=itm_logits_i,
                       mlm_logits=logits, cross=cross)
            if has_r:
                out.update(bbox_coord=coord)
        return losses

    @staticmethod
    def total_loss(losses, image_w=1.0, region_w=1.0, regions_use_bbox_only=False):
        """Pretrain.py:204-232: images.iter_perc * (itc + itm + mlm) + regions.iter_perc * (itc + itm + mlm + bbox + giou),
        or only the two box losses of the region iteration with config['regions_use_bbox_only'].  iter_perc defaults to 1
        (x2vlm_base_1b.yaml / x2vlm_large_1b.yaml set regions.iter_perc = 0.5)."""
        tot = image_w * sum(losses["image"].values())
        if losses["region"]:
            r = losses["region"]
            keys = ("loss_bbox", "loss_giou") if regions_use_bbox_only else tuple(r.keys())
            tot = tot + region_w * sum(r[k] for k in keys)
        return tot
