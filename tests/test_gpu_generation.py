"""GPU parity tests of the generation paths of the drop-in xbert (BertLMHeadModel, key/value cache,
history_states) against tests/golden/decoder_small.pt — outputs of the UNMODIFIED reference
(oracle/make_golden.py) — and against oracle/restate.py.  Tolerances as in test_gpu_modules.py (bf16 GEMM
operands, fp32 accumulation): logits rel-L2 <= 2e-2, losses 2e-2 relative, arg-max exact wherever the
reference's top-2 margin exceeds twice the observed logit error."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    return ((a.float().cpu() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def argmax_ok(got, ref):
    got = got.float().cpu()
    top2 = ref.topk(2, -1).values
    err = (got - ref).abs().max().item()
    sure = (top2[..., 0] - top2[..., 1]) > 2 * err
    return sure.any().item() and torch.equal(got.argmax(-1)[sure], ref.argmax(-1)[sure])


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "decoder_small.pt"))


def _decoder(g, dev, label_smoothing=0.0):
    from x2vlm_b200 import xbert
    c = g["cfg"]
    cfg = xbert.BertConfig(vocab_size=c["vocab_size"], hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                           num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                           max_position_embeddings=c["max_position_embeddings"], type_vocab_size=2, pad_token_id=0,
                           hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12)
    cfg.fusion_layer, cfg.encoder_width, cfg.embedding_dim = c["fusion_layer"], c["encoder_width"], c["hidden_size"]
    m = xbert.BertLMHeadModel(cfg, label_smoothing=label_smoothing)
    m.load_state_dict(g["state_dict"], strict=True)
    return m.to(dev).eval()


def test_lm_head_model_vqa_losses(dev, gold):
    g = gold
    m = _decoder(g, dev)
    t = lambda k: g[k].to(dev)
    with torch.no_grad():
        o = m(t("a_ids"), attention_mask=t("a_atts"), encoder_hidden_states=t("q_states"), encoder_attention_mask=t("q_atts"),
              labels=t("targets"), return_dict=True, reduction='none')
        m.label_smoothing = 0.1
        ls = m(t("a_ids"), attention_mask=t("a_atts"), encoder_hidden_states=t("q_states"), encoder_attention_mask=t("q_atts"),
               labels=t("targets"), return_dict=True, reduction='mean')
        # the loss-only call above is the fused vocabulary-GEMM + cross-entropy path (no logits are formed: o.logits is None);
        # the logits themselves come from the un-fused path, whose loss must agree with the fused one
        assert o.logits is None
        m.label_smoothing = 0.0
        m.config.x2k_fused_ce = False
        o2 = m(t("a_ids"), attention_mask=t("a_atts"), encoder_hidden_states=t("q_states"), encoder_attention_mask=t("q_atts"),
               labels=t("targets"), return_dict=True, reduction='none')
        m.config.x2k_fused_ce = True
    assert torch.allclose(o.loss, o2.loss, rtol=2e-3, atol=2e-3)
    o.logits = o2.logits
    valid = g["a_atts"].bool()
    assert rel_l2(o.logits.cpu()[valid], g["vqa_logits"][valid]) < 2e-2
    assert torch.allclose(o.loss.cpu(), g["vqa_loss"], rtol=2e-2, atol=2e-2)
    assert abs(float(ls.loss) - float(g["vqa_ls_loss"])) < 2e-2 * float(g["vqa_ls_loss"])
    assert argmax_ok(o.logits.cpu()[valid], g["vqa_logits"][valid])


def test_lm_head_model_trains_through_causal_path(dev, gold):
    """Training (labels given) uses the fused autograd layers with the causal 3-D mask; gradients vs the oracle."""
    from oracle import restate
    g = gold
    m = _decoder(g, dev)
    t = lambda k: g[k].to(dev)
    o = m(t("a_ids"), attention_mask=t("a_atts"), encoder_hidden_states=t("q_states"), encoder_attention_mask=t("q_atts"),
          labels=t("targets"), return_dict=True, reduction='none')   # eval mode: no dropout, deterministic
    o.loss.sum().backward()
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["state_dict"].items()}
    sd["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    sd["cls.predictions.decoder.bias"] = sd["cls.predictions.bias"]
    c = g["cfg"]
    h, _, _ = restate.bert_decoder(sd, "bert.", c["num_attention_heads"], c["num_hidden_layers"], g["a_ids"], g["a_atts"],
                                   g["q_states"], g["q_atts"])
    restate.lm_loss(restate.mlm_head(h, sd, "cls.predictions."), g["targets"], reduction="none").sum().backward()
    checked = 0
    for n, p in m.named_parameters():
        w = sd[n].grad
        if w is None or "key.bias" in n:
            continue
        e = rel_l2(p.grad, w)
        assert e < 3e-2, (n, e)
        checked += 1
    assert checked > 30


def test_key_value_cache_steps(dev, gold):
    g = gold
    m = _decoder(g, dev)
    t = lambda k: g[k].to(dev)
    ones = lambda n: torch.ones(4, n, dtype=torch.long, device=dev)
    kw = dict(encoder_hidden_states=t("q_states"), encoder_attention_mask=t("q_atts"), use_cache=True, return_dict=True)
    with torch.no_grad():
        s0 = m(t("a_ids")[:, :5], **kw)
        s1 = m(t("a_ids")[:, 5:6], attention_mask=ones(6), past_key_values=s0.past_key_values, **kw)
        s2 = m(t("a_ids")[:, 6:7], attention_mask=ones(7), past_key_values=s1.past_key_values, **kw)
    assert len(s1.past_key_values) == 2 and tuple(s1.past_key_values[0][0].shape) == (4, 2, 6, 64)
    assert rel_l2(s0.logits, g["st0_logits"]) < 2e-2
    assert rel_l2(s1.logits, g["st1_logits"]) < 2e-2
    assert rel_l2(s2.logits, g["st2_logits"]) < 2e-2
    assert rel_l2(s1.past_key_values[0][0], g["st1_k0"]) < 1e-2 and rel_l2(s1.past_key_values[1][1], g["st1_v1"]) < 1e-2
    assert argmax_ok(torch.cat((s0.logits, s1.logits, s2.logits), 1), g["full7_logits"])
    # the cached greedy loop reproduces the uncached arg-max continuation of the reference's 7-token pass
    with torch.no_grad():
        out = m.greedy_decode(t("a_ids")[:, :5], max_length=7, encoder_hidden_states=t("q_states"),
                              encoder_attention_mask=t("q_atts"))
    assert out.shape == (4, 7) and torch.equal(out[:, :5].cpu(), g["a_ids"][:, :5])
    top2 = g["st0_logits"][:, -1].topk(2, -1).values
    sure = (top2[:, 0] - top2[:, 1]) > 0.05
    assert torch.equal(out[:, 5].cpu()[sure], g["st0_logits"][:, -1].argmax(-1)[sure])


def test_history_states_captioning_loop(dev, gold):
    """model_generation.py:172-252 without the beam bookkeeping: [tokens..., MASK] per step, the cached layer
    inputs of the earlier positions passed as history_states, 3-D attention-mask slices."""
    g = gold
    m = _decoder(g, dev)
    t = lambda k: g[k].to(dev)
    Ltot = 12
    tril = torch.tril(torch.ones(Ltot, Ltot, dtype=torch.long, device=dev)).view(1, Ltot, Ltot).expand(4, Ltot, Ltot)
    pos = torch.arange(Ltot, device=dev).view(1, -1).expand(4, -1)
    tt = torch.zeros(4, Ltot, dtype=torch.long, device=dev)
    mask_tok = torch.full((4, 1), 103, dtype=torch.long, device=dev)
    a_ids = t("a_ids")
    curr, prev, next_pos = a_ids[:, :4], None, 4
    with torch.no_grad():
        for step in range(3):
            L = curr.shape[1]
            start = next_pos - L
            o = m.bert(torch.cat((curr, mask_tok), dim=1), attention_mask=tril[:, start:next_pos + 1, :next_pos + 1],
                       token_type_ids=tt[:, start:next_pos + 1], position_ids=pos[:, start:next_pos + 1],
                       encoder_hidden_states=t("img"), encoder_attention_mask=t("iatt"), output_hidden_states=True,
                       history_states=prev, is_decoder=True, return_dict=True)
            new = o.hidden_states
            assert len(new) == 3
            last = new[-1][:, -1:, :]
            assert rel_l2(last, g["hist_last"][step]) < 1e-2, step
            logits = m.cls(last)
            assert rel_l2(logits, g["hist_logits"][step]) < 2e-2, step
            assert argmax_ok(logits, g["hist_logits"][step])
            prev = [x[:, :-1, :] for x in new] if prev is None else [torch.cat((h, x[:, :-1, :]), dim=1) for h, x in zip(prev, new)]
            curr = a_ids[:, next_pos:next_pos + 1]
            next_pos += 1


def test_generation_paths_refuse_autograd(dev, gold):
    m = _decoder(gold, dev)
    ids = gold["a_ids"].to(dev)
    with torch.no_grad():
        s0 = m(ids[:, :5], use_cache=True, return_dict=True)
    with pytest.raises(NotImplementedError):
        m(ids[:, 5:6], past_key_values=s0.past_key_values, return_dict=True)
