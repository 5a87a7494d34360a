"""CPU tests: the oracle (oracle/restate.py, oracle/philox.py) against the committed golden fixtures that were
produced by running the unmodified reference (oracle/make_golden.py), plus published known-answer vectors."""
import os

import numpy as np
import torch

from oracle import philox, restate

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_raw(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_keep_scale_statistics():
    k = philox.keep_scale(1234, 77, 100003, 0.1)
    assert k.shape == (100003,)
    assert len(np.unique(k)) == 2 and k.min() == 0.0 and abs(float(k.max()) - 1.0 / 0.9) < 1e-4
    assert abs((k > 0).mean() - 0.9) < 0.01
    assert abs(float(k.mean()) - 1.0) < 0.01  # unbiased: E[keep * scale] = 1
    assert np.array_equal(k[:1000], philox.keep_scale(1234, 77, 1000, 0.1))  # prefix-stable


def test_restate_vision_matches_reference_golden():
    g = torch.load(os.path.join(GOLD, "vision_small.pt"))
    sd, c = g["state_dict"], g["cfg"]
    out = restate.vision_forward(g["image"], sd, "", c["depth"], c["num_heads"])
    assert torch.allclose(out, g["out_full"], atol=2e-5, rtol=1e-5)
    r, f = restate.vision_forward(g["region_image"], sd, "", c["depth"], c["num_heads"], g["idx_to_group_img"], g["image_atts"])
    assert torch.allclose(r, g["out_region"], atol=2e-5, rtol=1e-5)
    assert torch.allclose(f, g["out_region_full"], atol=2e-5, rtol=1e-5)
    assert torch.equal(restate.relative_position_index((14, 14)), sd["blocks.0.attn.relative_position_index"])


def test_restate_text_matches_reference_golden():
    g = torch.load(os.path.join(GOLD, "text_small.pt"))
    sd, c = g["state_dict"], g["cfg"]
    kw = dict(sd=sd, pfx="bert.", num_heads=c["num_attention_heads"], fusion_layer=c["fusion_layer"],
              num_layers=c["num_hidden_layers"])
    text = restate.bert_model(input_ids=g["ids"], attention_mask=g["atts"], mode="text", **kw)
    assert torch.allclose(text, g["text"], atol=2e-5, rtol=1e-5)
    cross = restate.bert_model(encoder_embeds=g["text"], attention_mask=g["atts"], enc_hidden=g["img"], enc_mask=g["iatt"],
                               mode="fusion", **kw)
    assert torch.allclose(cross, g["cross"], atol=2e-5, rtol=1e-5)
    seq = restate.bert_model(input_ids=g["ids"], attention_mask=g["atts"], enc_hidden=g["img"], enc_mask=g["iatt"],
                             mode="multi_modal", **kw)
    logits = restate.mlm_head(restate.gather_by_pos(seq, g["masked_pos"]), sd, "cls.predictions.")
    assert torch.allclose(logits, g["mlm_logits"], atol=5e-5, rtol=1e-5)
    loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), g["labels"].view(-1))
    assert abs(float(loss) - float(g["mlm_loss"])) < 1e-5
    t3 = restate.bert_model(input_ids=g["ids"], attention_mask=g["mask3d"], mode="text", **kw)
    assert torch.allclose(t3, g["text3d"], atol=2e-5, rtol=1e-5)
    # index/argmax surface is bit-exact
    assert torch.equal(logits.argmax(-1), g["mlm_logits"].argmax(-1))


def test_restate_decoder_matches_reference_golden():
    """Generation paths (BertLMHeadModel: causal mask, labels/reduction/label smoothing, key/value cache; captioning
    history_states loop) against the unmodified reference's outputs."""
    g = torch.load(os.path.join(GOLD, "decoder_small.pt"))
    sd, c = g["state_dict"], g["cfg"]
    kw = dict(sd=sd, pfx="bert.", num_heads=c["num_attention_heads"], num_layers=c["num_hidden_layers"])
    head = lambda h: restate.mlm_head(h, sd, "cls.predictions.")
    h, _, _ = restate.bert_decoder(input_ids=g["a_ids"], attention_mask=g["a_atts"], enc_hidden=g["q_states"],
                                   enc_mask=g["q_atts"], **kw)
    logits = head(h)
    assert torch.allclose(logits, g["vqa_logits"], atol=5e-5, rtol=1e-5)
    assert torch.allclose(restate.lm_loss(logits, g["targets"], reduction="none"), g["vqa_loss"], atol=1e-4, rtol=1e-5)
    assert abs(float(restate.lm_loss(logits, g["targets"], 0.1, "mean")) - float(g["vqa_ls_loss"])) < 1e-5
    # cache: prompt, then two single-token steps == the uncached 7-token pass
    h0, p0, _ = restate.bert_decoder(input_ids=g["a_ids"][:, :5], enc_hidden=g["q_states"], enc_mask=g["q_atts"], **kw)
    h1, p1, _ = restate.bert_decoder(input_ids=g["a_ids"][:, 5:6], enc_hidden=g["q_states"], enc_mask=g["q_atts"], past=p0, **kw)
    h2, _, _ = restate.bert_decoder(input_ids=g["a_ids"][:, 6:7], enc_hidden=g["q_states"], enc_mask=g["q_atts"], past=p1, **kw)
    assert torch.allclose(head(h0), g["st0_logits"], atol=5e-5, rtol=1e-5)
    assert torch.allclose(head(h1), g["st1_logits"], atol=5e-5, rtol=1e-5)
    assert torch.allclose(head(h2), g["st2_logits"], atol=5e-5, rtol=1e-5)
    assert torch.allclose(p1[0][0], g["st1_k0"], atol=2e-5) and torch.allclose(p1[1][1], g["st1_v1"], atol=2e-5)
    assert torch.allclose(torch.cat((head(h0), head(h1), head(h2)), 1), g["full7_logits"], atol=1e-4, rtol=1e-5)
    # captioning: history_states
    Ltot = 12
    tril = torch.tril(torch.ones(Ltot, Ltot, dtype=torch.long)).view(1, Ltot, Ltot).expand(4, Ltot, Ltot)
    pos = torch.arange(Ltot).view(1, -1).expand(4, -1)
    mask_tok = torch.full((4, 1), 103, dtype=torch.long)
    curr, prev, next_pos = g["a_ids"][:, :4], None, 4
    for step in range(3):
        L = curr.shape[1]
        start = next_pos - L
        _, _, hs = restate.bert_decoder(input_ids=torch.cat((curr, mask_tok), 1), attention_mask=tril[:, start:next_pos + 1, :next_pos + 1],
                                        position_ids=pos[:, start:next_pos + 1], enc_hidden=g["img"], enc_mask=g["iatt"],
                                        history=prev, **kw)
        assert torch.allclose(hs[-1][:, -1:], g["hist_last"][step], atol=2e-5, rtol=1e-5)
        assert torch.allclose(head(hs[-1][:, -1:]), g["hist_logits"][step], atol=5e-5, rtol=1e-5)
        prev = [x[:, :-1] for x in hs] if prev is None else [torch.cat((p, x[:, :-1]), 1) for p, x in zip(prev, hs)]
        curr = g["a_ids"][:, next_pos:next_pos + 1]
        next_pos += 1
