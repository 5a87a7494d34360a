"""Pin oracle/restate.py against the UNMODIFIED reference (run through oracle/ref_shim.py).

The reference ships no tests/golden vectors, so its own execution is the pin.
These tests only run where /root/reference exists (build container)."""
import os

import pytest
import torch

from oracle import ref_shim, restate
from x2vlm_b200 import synth

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    m = ref_shim.build_reference_xvlm(seed=0)
    # exercise paths whose parameters initialise to constants (rel-pos table = 0, gamma = 0.1)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif "gamma_" in n:
                p.copy_(0.1 + torch.randn(p.shape, generator=g) * 0.05)
            elif n.endswith(".bias") or "q_bias" in n or "v_bias" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    m.eval()
    return m


def _sd(m):
    return {k: v.detach() for k, v in m.state_dict().items()}


def test_relative_position_index(ref):
    idx = restate.relative_position_index((14, 14))
    assert torch.equal(idx, ref.vision_encoder.blocks[0].attn.relative_position_index)


def test_beit_block_and_vision(ref):
    sd = _sd(ref)
    b = synth.image_text_batch(2, 30, seed=1)
    with torch.no_grad():
        x = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(3))
        y_ref, p_ref = ref.vision_encoder.blocks[3](x)
        y, p = restate.beit_block(x, sd, "vision_encoder.blocks.3.", 12)
        assert torch.allclose(y, y_ref, atol=1e-5, rtol=1e-5)
        assert torch.allclose(p, p_ref, atol=1e-6)
        e_ref = ref.vision_encoder(b["image"])
        e = restate.vision_forward(b["image"], sd, "vision_encoder.", 12, 12)
        assert torch.allclose(e, e_ref, atol=2e-4, rtol=1e-4)


def test_vision_region_mode(ref):
    sd = _sd(ref)
    rb = synth.region_batch(3, 8, 30, seed=2)
    with torch.no_grad():
        r_ref, f_ref = ref.vision_encoder(rb["image"], idx_to_group_img=rb["idx_to_group_img"], image_atts=rb["image_atts"])
        r, f = restate.vision_forward(rb["image"], sd, "vision_encoder.", 12, 12, rb["idx_to_group_img"], rb["image_atts"])
    assert torch.allclose(r, r_ref, atol=2e-4, rtol=1e-4) and torch.allclose(f, f_ref, atol=2e-4, rtol=1e-4)


def test_text_and_fusion(ref):
    sd = _sd(ref)
    b = synth.image_text_batch(2, 30, seed=1)
    atts = b["text_atts"].clone()
    atts[1, 20:] = 0  # padded caption
    kw = dict(sd=sd, pfx="text_encoder.bert.", num_heads=12, fusion_layer=12, num_layers=18)
    with torch.no_grad():
        ie, ia = ref.get_vision_embeds(b["image"])
        te_ref = ref.get_text_embeds(b["text_ids"], atts)
        te = restate.bert_model(input_ids=b["text_ids"], attention_mask=atts, mode="text", **kw)
        assert torch.allclose(te, te_ref, atol=2e-4, rtol=1e-4)
        ce_ref = ref.get_cross_embeds(ie, ia, text_embeds=te_ref, text_atts=atts)
        ce = restate.bert_model(encoder_embeds=te_ref, attention_mask=atts, enc_hidden=ie, enc_mask=ia, mode="fusion", **kw)
        assert torch.allclose(ce, ce_ref, atol=3e-4, rtol=1e-4)
        # 3-D self-attention mask (captioning layout, models/model_generation.py:122-123)
        m3 = torch.tril(torch.ones(30, 30)).unsqueeze(0).expand(2, -1, -1).contiguous()
        o_ref = ref.text_encoder.bert(b["text_ids"], attention_mask=m3, mode="text", return_dict=True).last_hidden_state
        o = restate.bert_model(input_ids=b["text_ids"], attention_mask=m3, mode="text", **kw)
        assert torch.allclose(o, o_ref, atol=2e-4, rtol=1e-4)


def test_pretrain_losses_match_reference(ref, monkeypatch):
    sd = _sd(ref)
    B = 4
    b = synth.image_text_batch(B, 30, seed=11)
    ineg, tneg = synth.hard_negative_indices(B, seed=3)
    monkeypatch.setattr(ref, "get_hard_negatives", lambda *a, **k: (ineg.tolist(), tneg.tolist()))
    with torch.no_grad():
        l_ref = ref(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"],
                    masked_pos=b["masked_pos"], masked_ids=b["masked_ids"])
        l = restate.pretrain_forward(sd, restate.Shapes(), b["image"], b["text_ids"], b["text_atts"], b["text_ids_masked"],
                                     b["masked_pos"], b["masked_ids"], ineg, tneg)
    for k in ("loss_itc", "loss_itm", "loss_mlm"):
        assert abs(float(l[k]) - float(l_ref[k])) < 2e-4 * max(1.0, abs(float(l_ref[k]))), k


def test_pretrain_bbox_losses_match_reference(ref, monkeypatch):
    sd = _sd(ref)
    rb = synth.region_batch(3, 6, 30, seed=21)
    ineg, tneg = synth.hard_negative_indices(6, seed=4)
    monkeypatch.setattr(ref, "get_hard_negatives", lambda *a, **k: (ineg.tolist(), tneg.tolist()))
    with torch.no_grad():
        l_ref = ref(rb["image"], rb["text_ids"], rb["text_atts"], text_ids_masked=rb["text_ids_masked"],
                    masked_pos=rb["masked_pos"], masked_ids=rb["masked_ids"], image_atts=rb["image_atts"],
                    idx_to_group_img=rb["idx_to_group_img"], target_bbox=rb["target_bbox"], is_image=rb["is_image"],
                    ret_bbox_loss=True)
        l = restate.pretrain_forward(sd, restate.Shapes(), rb["image"], rb["text_ids"], rb["text_atts"], rb["text_ids_masked"],
                                     rb["masked_pos"], rb["masked_ids"], ineg, tneg, image_atts=rb["image_atts"],
                                     idx_to_group_img=rb["idx_to_group_img"], target_bbox=rb["target_bbox"],
                                     is_image=rb["is_image"], ret_bbox_loss=True)
    for k in ("loss_itc", "loss_itm", "loss_mlm", "loss_bbox", "loss_giou"):
        assert abs(float(l[k]) - float(l_ref[k])) < 2e-4 * max(1.0, abs(float(l_ref[k]))), k


def test_xbert_drop_path_matches_reference():
    """xbert's per-POSITION DropPath (models/xbert.py:518-548; only configs/finetune/refcoco_grounding_large.yaml turns
    it on): the reference layer in train mode and the oracle consume the same three torch.rand((1, L, 1)) draws."""
    ref_shim.install()
    from models import xbert as rxbert
    cfg = rxbert.BertConfig(vocab_size=64, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256,
                            max_position_embeddings=32, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.0)
    cfg.fusion_layer, cfg.encoder_width = 1, 128
    cfg.text_drop_path_rate, cfg.cross_drop_path_rate = 0.4, 0.4
    torch.manual_seed(3)
    enc = rxbert.BertEncoder(cfg)
    assert cfg.hidden_dropout_prob == 0.0                      # the reference zeroes it when drop path is on (:637-640)
    layer = enc.layer[2]                                        # fusion layer with the full cross rate
    assert abs(layer.output.drop_path.drop_prob - 0.4) < 1e-6
    sd = {"l." + k: v.detach() for k, v in layer.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    B, L = 3, 11
    hidden, img = torch.randn(B, L, 128, generator=g), torch.randn(B, 17, 128, generator=g)
    layer.train()
    torch.manual_seed(77)
    with torch.no_grad():
        want = layer(hidden, None, None, img, None)[0]
    torch.manual_seed(77)
    dp = {k: restate.drop_path_scale(L, 0.4) for k in ("self", "cross", "ffn")}   # same order as the reference's draws
    assert any((v == 0).any() for v in dp.values()) and any((v > 0).any() for v in dp.values())
    with torch.no_grad():
        got = restate.bert_layer(hidden, None, sd, "l.", 2, enc_hidden=img, cross_mask=None, train=True, p_attn=0.0, p_hidden=0.0,
                                 dp_scales=dp)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_interpolate_pos_embed_same_resolution_matches_reference(ref):
    """Checkpoint adaptation helper (models/beit2.py:664-754) on the reference's own vision state_dict at its own
    resolution: same keys dropped, every tensor untouched.  (The resizing branch of the reference cannot run here:
    SciPy >= 1.14 removed interp2d; ours uses the documented RectBivariateSpline replacement, tests/test_host_logic.py.)"""
    from models import beit2 as rbeit
    from x2vlm_b200 import beit2
    sd = {k: v.clone() for k, v in ref.vision_encoder.state_dict().items()}
    want = rbeit.interpolate_pos_embed(ref.vision_encoder, {k: v.clone() for k, v in sd.items()})
    ours = beit2.beit_base_patch16(img_size=224, drop_rate=0.0, drop_path_rate=0.1, attn_drop_rate=0.0, use_mean_pooling=True,
                                   init_scale=0.001, use_rel_pos_bias=True, use_abs_pos_emb=False, init_values=0.1, qkv_bias=True)
    got = beit2.interpolate_pos_embed(ours, {k: v.clone() for k, v in sd.items()})
    assert sorted(got) == sorted(want) and not any("relative_position_index" in k for k in got)
    assert all(torch.equal(got[k], want[k]) for k in want)
    missing, unexpected, ignored = beit2.load_state_dict(ours, got)
    assert missing == [] and unexpected == [] and len(ignored) == 12


def test_top_k_top_p_filtering_matches_reference():
    ref_shim.install()
    from models import xbert as rxbert
    from x2vlm_b200 import xbert
    g = torch.Generator().manual_seed(0)
    for top_k, top_p, keep in ((0, 0.9, 1), (5, 1.0, 1), (7, 0.6, 3), (0, 0.3, 2), (50, 0.95, 1)):
        logits = torch.randn(6, 40, generator=g) * 3
        want = rxbert.top_k_top_p_filtering(logits.clone(), top_k=top_k, top_p=top_p, min_tokens_to_keep=keep)
        got = xbert.top_k_top_p_filtering(logits.clone(), top_k=top_k, top_p=top_p, min_tokens_to_keep=keep)
        assert torch.equal(got, want), (top_k, top_p, keep)


def test_itc_idx_labels_and_negative_weights_match_reference(ref, monkeypatch):
    """Retrieval fine-tuning passes sample ids (models/model_retrieval.py:13-24): soft ITC labels over equal ids
    (models/xvlm.py:811-824) and hard-negative weights with every same-id pair zeroed (:838-846)."""
    from types import SimpleNamespace
    from x2vlm_b200 import pretrain
    g = torch.Generator().manual_seed(2)
    fi = torch.nn.functional.normalize(torch.randn(6, 256, generator=g), dim=-1)
    ft = torch.nn.functional.normalize(torch.randn(6, 256, generator=g), dim=-1)
    idx = torch.tensor([3, 7, 3, 9, 7, 1])
    me = SimpleNamespace(temp=ref.temp.detach().clone())
    me.hard_negative_weights = lambda *a, **k: pretrain.XVLM.hard_negative_weights(me, *a, **k)
    with torch.no_grad():
        for ids in (None, idx):
            want = ref.get_contrastive_loss(fi, ft, idx=ids)
            got = pretrain.XVLM.get_contrastive_loss(me, fi, ft, idx=ids)
            assert abs(float(got) - float(want)) < 1e-6, ids
    seen = []
    real = torch.multinomial
    monkeypatch.setattr(torch, "multinomial", lambda w, n, *a, **k: (seen.append(w.clone()), real(w, n, *a, **k))[1])
    ref.get_hard_negatives(fi, ft, idx=idx)           # B draws from weights_t2i rows, then B from weights_i2t rows
    monkeypatch.undo()
    w_t2i_ref, w_i2t_ref = torch.stack(seen[:6]), torch.stack(seen[6:12])
    w_i2t, w_t2i = pretrain.XVLM.hard_negative_weights(me, fi, ft, idx=idx)
    assert torch.allclose(w_i2t, w_i2t_ref, atol=1e-7) and torch.allclose(w_t2i, w_t2i_ref, atol=1e-7)
    assert float(w_i2t[0, 2]) == 0.0 and float(w_i2t[1, 4]) == 0.0 and float(w_i2t[0, 1]) > 0.0
    ineg, tneg = pretrain.XVLM.get_hard_negatives(me, fi, ft, idx=idx)
    assert all(int(idx[i]) != int(idx[j]) for i, j in enumerate(ineg.tolist())) and ineg.shape == (6,) and tneg.shape == (6,)


def test_retrieval_scores_match_reference_evaluation(ref):
    """oracle.restate.retrieval_scores against the reference's own Retrieval.evaluation loop (Retrieval.py:71-157) run
    on stub loader / tokenizer objects: ITC similarities, top-k candidates and ITM re-rank scores, both directions."""
    n_img, n_txt, k = 3, 5, 2
    b = synth.image_text_batch(n_txt, 40, seed=7)
    images, ids, atts = b["image"][:n_img], b["text_ids"], b["text_atts"].clone()
    atts[1, 30:] = 0
    s_i2t, s_t2i = ref_shim.run_reference_retrieval(ref, images, ids, atts, k, "cpu", image_bs=2, text_bs=2)
    sd = _sd(ref)
    with torch.no_grad():
        ie = restate.vision_forward(images, sd, "vision_encoder.", 12, 12)
        te = restate.bert_model(sd, "text_encoder.bert.", 12, 12, 18, input_ids=ids, attention_mask=atts, mode="text")
        fi, ft = restate.get_features(ie, te, sd)
        w_i2t, w_t2i = restate.retrieval_scores(sd, restate.Shapes(), ie, te, atts, fi @ ft.t(), k)
    assert torch.equal(torch.from_numpy(s_i2t) > -99, w_i2t > -99) and torch.equal(torch.from_numpy(s_t2i) > -99, w_t2i > -99)
    assert (w_i2t - torch.from_numpy(s_i2t)).abs().max() < 1e-4 and (w_t2i - torch.from_numpy(s_t2i)).abs().max() < 1e-4


def test_video_avgpool_matches_reference():
    """oracle.restate.video_forward against the reference's XVLMBase.get_vision_embeds on a 5-D input
    (video_encoding='avgpool', learned per-frame offsets; models/xvlm.py:615-661)."""
    m = ref_shim.build_reference_xvlm(ref_shim.base_config(video_encoding="avgpool", frame_len=3, add_frame_pos=True,
                                                           vision_num_hidden_layers=2), seed=1)
    m.eval()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        m.absolute_frame_pos_embed.copy_(torch.randn(m.absolute_frame_pos_embed.shape, generator=g) * 0.1)
    frames = torch.randn(2, 3, 3, 224, 224, generator=g)
    sd = _sd(m)
    with torch.no_grad():
        want, want_atts = m.get_vision_embeds(frames)
        got = restate.video_forward(frames, sd, restate.Shapes(vision_depth=2), sd["absolute_frame_pos_embed"])
    assert want.shape == (2, 197, 768) and want_atts.shape == (2, 197)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


def test_beit_drop_path_two_independent_draws(ref):
    """The reference calls self.drop_path twice per block (models/beit2.py:204-207): the attention and the MLP branch get
    independent per-sample masks.  Replaying the two torch.rand draws of timm's drop_path through the oracle reproduces
    the reference's train-mode block output — the contract the fused block's (dp_scale, dp_scale2) pair implements."""
    sd = _sd(ref)
    blk = ref.vision_encoder.blocks[7]
    p = blk.drop_path.drop_prob
    assert p > 0
    x = torch.randn(6, 197, 768, generator=torch.Generator().manual_seed(11))
    blk.train()
    try:
        torch.manual_seed(123)
        with torch.no_grad():
            y_ref, _ = blk(x)
    finally:
        blk.eval()
    torch.manual_seed(123)
    keep = 1 - p
    dp1 = torch.floor(keep + torch.rand(6, 1, 1)).view(-1) / keep
    dp2 = torch.floor(keep + torch.rand(6, 1, 1)).view(-1) / keep
    with torch.no_grad():
        y, _ = restate.beit_block(x, sd, "vision_encoder.blocks.7.", 12, dp1, dp2)
    assert torch.allclose(y, y_ref, atol=1e-5, rtol=1e-5)


def _itc_worker(rank, world, port, q):
    """One rank of the ITC loss under data parallelism: ONE packed all-gather (pretrain.allgather_packed, what the mixed
    step does) against the reference's own XVLMBase.get_contrastive_loss with its two all-gathers per loss."""
    import types
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref_shim.install()
        import models.xvlm as ref_xvlm
        from x2vlm_b200 import pretrain
        g = torch.Generator().manual_seed(50 + rank)
        D = 8
        feats = [torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=-1).requires_grad_(True) for n in (4, 4, 3, 3)]
        stub = types.SimpleNamespace(temp=torch.tensor(0.07), embed_dim=D)
        # ours: one collective for the four matrices, then the two losses
        gathered = pretrain.allgather_packed(feats)
        ours = (pretrain.XVLM.get_contrastive_loss(stub, feats[0], feats[1], gathered=(gathered[0], gathered[1]))
                + pretrain.XVLM.get_contrastive_loss(stub, feats[2], feats[3], gathered=(gathered[2], gathered[3])))
        g_ours = torch.autograd.grad(ours, feats)
        # the reference's method (models/xvlm.py:794-826), unmodified, on the same per-rank features
        ref = (ref_xvlm.XVLMBase.get_contrastive_loss(stub, feats[0], feats[1])
               + ref_xvlm.XVLMBase.get_contrastive_loss(stub, feats[2], feats[3]))
        g_ref = torch.autograd.grad(ref, feats)
        ok = torch.allclose(ours, ref, atol=1e-6) and all(torch.allclose(a, b, atol=1e-6) for a, b in zip(g_ours, g_ref))
        # rank-major layout of the gathered matrices: rows [r * B, (r + 1) * B) hold rank r's features
        mine = gathered[2][rank * 3:(rank + 1) * 3]
        ok = ok and torch.equal(mine, feats[2].detach())
        q.put((rank, bool(ok), float(ours), float(ref)))
    finally:
        dist.destroy_process_group()


def test_itc_packed_allgather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_itc_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(30) for p in procs]
    assert all(ok for _, ok, _, _ in res), res
    assert abs(res[0][2] - res[1][2]) < 1e-6  # the loss is a function of the gathered batch: identical on both ranks
