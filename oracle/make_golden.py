"""TEST INFRASTRUCTURE — generate tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE (through
oracle/ref_shim.py) on small seeded models and inputs.  Run in the build container only:

    python -m oracle.make_golden

The reference has no golden vectors of its own; these fixtures travel to the GPU box, where
/root/reference does not exist, and pin both oracle/restate.py and the CUDA path against the reference.
Small widths keep the files a few MB; head_dim stays 64 (the kernels' head size).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from x2vlm_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def perturb(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif "gamma_" in n:
                p.copy_(0.1 + torch.randn(p.shape, generator=g) * 0.05)
            elif n.endswith("bias") and p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "LayerNorm.weight" in n or "norm1.weight" in n or "norm2.weight" in n or "fc_norm.weight" in n:
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)


def main():
    ref_shim.install()
    ref_shim.init_dist()
    os.makedirs(OUT, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(ref_shim.workdir())
    try:
        from functools import partial
        import torch.nn as nn
        from models import beit2 as rbeit
        from models import xbert as rxbert
        # ---- vision: 2 blocks, width 128 (2 heads of 64), 224 px ----
        torch.manual_seed(0)
        vis = rbeit.VisionTransformer(img_size=224, patch_size=16, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_rate=0.0, drop_path_rate=0.1,
                                      attn_drop_rate=0.0, use_mean_pooling=True, init_scale=0.001, use_rel_pos_bias=True,
                                      use_abs_pos_emb=False, init_values=0.1, qkv_bias=True)
        perturb(vis, 1)
        vis.eval()
        b = synth.image_text_batch(3, 24, seed=5)
        rb = synth.region_batch(2, 5, 24, seed=6)
        with torch.no_grad():
            out_full = vis(b["image"])
            out_region, out_region_full = vis(rb["image"], idx_to_group_img=rb["idx_to_group_img"], image_atts=rb["image_atts"])
        torch.save({"state_dict": {k: v.clone() for k, v in vis.state_dict().items()}, "image": b["image"],
                    "region_image": rb["image"], "idx_to_group_img": rb["idx_to_group_img"], "image_atts": rb["image_atts"],
                    "out_full": out_full, "out_region": out_region, "out_region_full": out_region_full,
                    "cfg": dict(embed_dim=128, depth=2, num_heads=2)}, os.path.join(OUT, "vision_small.pt"))
        # ---- text + fusion + MLM head: 3 layers (fusion from layer 2), width 128, vocab 1024 ----
        cfg = rxbert.BertConfig(vocab_size=1024, hidden_size=128, num_hidden_layers=3, num_attention_heads=2,
                                intermediate_size=512, max_position_embeddings=64, type_vocab_size=2, pad_token_id=0,
                                hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12)
        cfg.fusion_layer, cfg.encoder_width, cfg.embedding_dim = 2, 128, 128
        torch.manual_seed(1)
        mlm = rxbert.BertForMaskedLM(cfg)
        perturb(mlm, 2)
        mlm.eval()
        g = torch.Generator().manual_seed(9)
        ids = torch.randint(5, 1024, (3, 24), generator=g)
        atts = torch.ones(3, 24, dtype=torch.long)
        atts[2, 17:] = 0
        masked_pos = torch.stack([torch.randperm(22, generator=g)[:6].sort().values + 1 for _ in range(3)])
        labels = torch.gather(ids, 1, masked_pos)
        img = torch.randn(3, 197, 128, generator=g)
        iatt = torch.ones(3, 197, dtype=torch.long)
        iatt[1, 100:] = 0
        with torch.no_grad():
            text = mlm.bert(ids, attention_mask=atts, return_dict=True, mode="text").last_hidden_state
            cross = mlm.bert(encoder_embeds=text, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=iatt,
                             return_dict=True, mode="fusion").last_hidden_state
            o = mlm(ids, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=iatt, return_dict=True,
                    labels=labels, masked_pos=masked_pos)
            m3 = torch.tril(torch.ones(24, 24)).unsqueeze(0).expand(3, -1, -1).contiguous()
            text3d = mlm.bert(ids, attention_mask=m3, return_dict=True, mode="text").last_hidden_state
        torch.save({"state_dict": {k: v.clone() for k, v in mlm.state_dict().items()}, "ids": ids, "atts": atts,
                    "masked_pos": masked_pos, "labels": labels, "img": img, "iatt": iatt, "text": text, "cross": cross,
                    "mlm_logits": o.logits, "mlm_loss": o.loss, "mask3d": m3, "text3d": text3d,
                    "cfg": dict(vocab_size=1024, hidden_size=128, num_hidden_layers=3, num_attention_heads=2,
                                intermediate_size=512, max_position_embeddings=64, fusion_layer=2, encoder_width=128)},
                   os.path.join(OUT, "text_small.pt"))
        # ---- generation paths: BertLMHeadModel (VQA answer decoder) and the captioning history_states loop ----
        dcfg = rxbert.BertConfig(vocab_size=1024, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                 intermediate_size=512, max_position_embeddings=64, type_vocab_size=2, pad_token_id=0,
                                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12)
        dcfg.fusion_layer, dcfg.encoder_width, dcfg.embedding_dim = 0, 128, 128
        torch.manual_seed(3)
        dec = rxbert.BertLMHeadModel(dcfg, label_smoothing=0.0)
        perturb(dec, 4)
        dec.eval()
        g = torch.Generator().manual_seed(11)
        a_ids = torch.randint(5, 1024, (4, 9), generator=g)
        a_atts = torch.ones(4, 9, dtype=torch.long)
        a_atts[1, 6:] = 0
        a_atts[3, 4:] = 0
        targets = a_ids.masked_fill(a_atts == 0, -100)
        q_states = torch.randn(4, 21, 128, generator=g)
        q_atts = torch.ones(4, 21, dtype=torch.long)
        q_atts[2, 15:] = 0
        with torch.no_grad():
            vqa = dec(a_ids, attention_mask=a_atts, encoder_hidden_states=q_states, encoder_attention_mask=q_atts,
                      labels=targets, return_dict=True, reduction='none')
            dec.label_smoothing = 0.1
            vqa_ls = dec(a_ids, attention_mask=a_atts, encoder_hidden_states=q_states, encoder_attention_mask=q_atts,
                         labels=targets, return_dict=True, reduction='mean')
            dec.label_smoothing = 0.0
            # HF-style cache: prompt of 5 tokens fills the cache, then two single-token steps
            st0 = dec(a_ids[:, :5], encoder_hidden_states=q_states, encoder_attention_mask=q_atts, use_cache=True,
                      return_dict=True)
            st1 = dec(a_ids[:, 5:6], attention_mask=torch.ones(4, 6, dtype=torch.long), encoder_hidden_states=q_states,
                      encoder_attention_mask=q_atts, past_key_values=st0.past_key_values, use_cache=True, return_dict=True)
            st2 = dec(a_ids[:, 6:7], attention_mask=torch.ones(4, 7, dtype=torch.long), encoder_hidden_states=q_states,
                      encoder_attention_mask=q_atts, past_key_values=st1.past_key_values, use_cache=True, return_dict=True)
            full7 = dec(a_ids[:, :7], encoder_hidden_states=q_states, encoder_attention_mask=q_atts, use_cache=False,
                        return_dict=True)
            # captioning loop (model_generation.py:172-252, beam bookkeeping left out): [tokens..., MASK] with the
            # cached layer INPUTS of the previous positions as history_states
            Ltot = 12
            tril = torch.tril(torch.ones(Ltot, Ltot, dtype=torch.long)).view(1, Ltot, Ltot).expand(4, Ltot, Ltot)
            pos = torch.arange(Ltot).view(1, -1).expand(4, -1)
            tt = torch.zeros(4, Ltot, dtype=torch.long)
            img = torch.randn(4, 197, 128, generator=g)
            iatt = torch.ones(4, 197, dtype=torch.long)
            mask_tok = torch.full((4, 1), 103, dtype=torch.long)
            curr, prev, next_pos = a_ids[:, :4], None, 4
            hist_logits, hist_last = [], []
            for step in range(3):
                L = curr.shape[1]
                start = next_pos - L
                x_ids = torch.cat((curr, mask_tok), dim=1)
                o = dec.bert(x_ids, attention_mask=tril[:, start:next_pos + 1, :next_pos + 1],
                             token_type_ids=tt[:, start:next_pos + 1], position_ids=pos[:, start:next_pos + 1],
                             encoder_hidden_states=img, encoder_attention_mask=iatt, output_hidden_states=True,
                             history_states=prev, is_decoder=True, return_dict=True)
                new = o.hidden_states
                hist_last.append(new[-1][:, -1:, :].clone())
                hist_logits.append(dec.cls(new[-1][:, -1:, :]).clone())
                prev = [x[:, :-1, :] for x in new] if prev is None else \
                    [torch.cat((h, x[:, :-1, :]), dim=1) for h, x in zip(prev, new)]
                curr = a_ids[:, next_pos:next_pos + 1]
                next_pos += 1
        torch.save({"state_dict": {k: v.clone() for k, v in dec.state_dict().items()}, "a_ids": a_ids, "a_atts": a_atts,
                    "targets": targets, "q_states": q_states, "q_atts": q_atts, "vqa_loss": vqa.loss, "vqa_logits": vqa.logits,
                    "vqa_ls_loss": vqa_ls.loss, "st0_logits": st0.logits, "st1_logits": st1.logits, "st2_logits": st2.logits,
                    "st1_k0": st1.past_key_values[0][0], "st1_v1": st1.past_key_values[1][1], "full7_logits": full7.logits,
                    "img": img, "iatt": iatt, "hist_logits": torch.stack(hist_logits), "hist_last": torch.stack(hist_last),
                    "cfg": dict(vocab_size=1024, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                intermediate_size=512, max_position_embeddings=64, fusion_layer=0, encoder_width=128)},
                   os.path.join(OUT, "decoder_small.pt"))
    finally:
        os.chdir(cwd)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
