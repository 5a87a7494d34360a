"""GPU parity tests of the two caller-side workloads BASELINE.json lists next to pre-training: retrieval scoring
(ITC all-pairs + top-k ITM re-rank, Retrieval.py:71-157) and video input (frames -> avgpool, models/xvlm.py:615-661),
against oracle/restate.py on the same weights.  Index work (top-k candidate sets from identical similarities) is
bit-exact; scores carry the bf16 tolerance of test_gpu_modules.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.float().cpu() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _perturbed(m, seed):
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "relative_position_bias_table" in n:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.5)
            elif "gamma_" in n:
                p.copy_(0.1 + torch.randn(p.shape, generator=gen) * 0.05)
    return m


def test_retrieval_rerank_vs_oracle(dev):
    from oracle import restate
    from x2vlm_b200 import pretrain, retrieval, synth
    torch.manual_seed(0)
    cfg = pretrain.base_config(vision_num_hidden_layers=2, text_num_hidden_layers=4, text_fusion_start_at=2)
    m = _perturbed(pretrain.XVLM(cfg), 7)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev).eval()
    shp = restate.Shapes(vision_depth=2, vision_heads=12, text_heads=12, fusion_layer=2, num_layers=4)
    n_img, n_txt, k = 5, 11, 4
    b = synth.image_text_batch(n_txt, 40, seed=77)
    images, ids, atts = b["image"][:n_img], b["text_ids"], b["text_atts"].clone()
    atts[3, 25:] = 0
    atts[8, 31:] = 0
    s_i2t, s_t2i, sims = retrieval.evaluation(m, images.to(dev), ids.to(dev), atts.to(dev), k_test=k, image_bs=2, text_bs=4,
                                              rows_per_call=3)
    with torch.no_grad():
        ie = restate.vision_forward(images, sd, "vision_encoder.", 2, 12)
        te = restate.bert_model(sd, "text_encoder.bert.", 12, 2, 4, input_ids=ids, attention_mask=atts, mode="text")
        fi, ft = restate.get_features(ie, te, sd)
        want_sims = fi @ ft.t()
        assert (sims.cpu() - want_sims).abs().max() < 2e-2
        # candidate sets from IDENTICAL similarities are identical (index work is bit-exact)
        assert torch.equal(sims.topk(k, dim=1).indices.cpu(), sims.cpu().topk(k, dim=1).indices)
        top_i2t = [(s_i2t[i] > -99).nonzero().flatten().cpu() for i in range(n_img)]
        top_t2i = [(s_t2i[j] > -99).nonzero().flatten().cpu() for j in range(n_txt)]
        assert all(len(t) == k for t in top_i2t) and all(len(t) == k for t in top_t2i)
        for i in range(n_img):
            assert set(top_i2t[i].tolist()) == set(sims[i].topk(k).indices.tolist())
        for j in range(n_txt):
            assert set(top_t2i[j].tolist()) == set(sims[:, j].topk(k).indices.tolist())
        w_i2t, w_t2i = restate.retrieval_scores(sd, shp, ie, te, atts, want_sims, k, top_i2t, top_t2i)
    assert torch.equal(s_i2t.cpu() > -99, w_i2t > -99) and torch.equal(s_t2i.cpu() > -99, w_t2i > -99)
    mi, mt = w_i2t > -99, w_t2i > -99
    scale = max(1.0, float(w_i2t[mi].abs().max()))
    assert (s_i2t.cpu()[mi] - w_i2t[mi]).abs().max() < 3e-2 * scale
    assert (s_t2i.cpu()[mt] - w_t2i[mt]).abs().max() < 3e-2 * scale
    # recall bookkeeping (Retrieval.py:itm_eval) on a trivial ground truth runs and is bounded
    r = retrieval.itm_eval(s_i2t, s_t2i, txt2img={j: j % n_img for j in range(n_txt)},
                           img2txt={i: [j for j in range(n_txt) if j % n_img == i] for i in range(n_img)})
    assert 0.0 <= r["r_mean"] <= 100.0


def test_video_avgpool_vs_oracle(dev):
    from oracle import restate
    from x2vlm_b200 import pretrain
    torch.manual_seed(1)
    cfg = pretrain.base_config(vision_num_hidden_layers=2, text_num_hidden_layers=2, text_fusion_start_at=1,
                               video_encoding="avgpool", frame_len=4, add_frame_pos=True)
    m = _perturbed(pretrain.XVLM(cfg), 9)
    assert "absolute_frame_pos_embed" in m.state_dict() and m.absolute_frame_pos_embed.shape == (1, 4, 1, 768)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev).eval()
    frames = torch.randn(2, 4, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        emb, att = m.get_vision_embeds(frames.to(dev))
        want = restate.video_forward(frames, sd, restate.Shapes(vision_depth=2), sd["absolute_frame_pos_embed"])
    assert emb.shape == (2, 197, 768) and att.shape == (2, 197) and att.dtype == torch.long
    assert rel_l2(emb, want) < 1e-2
    # gradients reach the per-frame offset and the encoder through the time mean
    m.train()
    e, _ = m.get_vision_embeds(frames.to(dev))
    e.square().mean().backward()
    assert m.absolute_frame_pos_embed.grad is not None and float(m.absolute_frame_pos_embed.grad.abs().sum()) > 0
    assert m.vision_encoder.blocks[0].mlp.fc1.weight.grad is not None


def test_large_width_step_vs_oracle(dev):
    """BASELINE config 3 shapes (beit2-large / bert-large: width 1024, 16 heads, MLP 4096) at reduced depth: losses of
    the mixed image + region step against the oracle, and one optimizer-free backward."""
    from oracle import restate
    from x2vlm_b200 import pretrain, synth
    torch.manual_seed(2)
    cfg = pretrain.large_config(vision_num_hidden_layers=2, text_num_hidden_layers=3, text_fusion_start_at=2)
    m = _perturbed(pretrain.XVLM(cfg), 11)
    assert m.vision_width == 1024 and m.text_width == 1024
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev).eval()
    shp = restate.Shapes(vision_depth=2, vision_heads=16, text_heads=16, fusion_layer=2, num_layers=3)
    ib, rb = synth.image_text_batch(3, 40, seed=51), synth.region_batch(2, 3, 40, seed=52)
    neg_i, neg_r = synth.hard_negative_indices(3, 1), synth.hard_negative_indices(3, 2)
    with torch.no_grad():
        want_i = restate.pretrain_forward(sd, shp, ib["image"], ib["text_ids"], ib["text_atts"], ib["text_ids_masked"],
                                          ib["masked_pos"], ib["masked_ids"], *neg_i)
        want_r = restate.pretrain_forward(sd, shp, rb["image"], rb["text_ids"], rb["text_atts"], rb["text_ids_masked"],
                                          rb["masked_pos"], rb["masked_ids"], *neg_r, image_atts=rb["image_atts"],
                                          idx_to_group_img=rb["idx_to_group_img"], target_bbox=rb["target_bbox"],
                                          is_image=rb["is_image"], ret_bbox_loss=True)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    got = m.forward_mixed(to(ib), to(rb), tuple(t.to(dev) for t in neg_i), tuple(t.to(dev) for t in neg_r))
    for name, want in (("image", want_i), ("region", want_r)):
        for k, w in want.items():
            assert abs(float(got[name][k]) - float(w)) < 2e-2 * max(1.0, abs(float(w))), (name, k, float(got[name][k]), float(w))
    m.total_loss(got).backward()
    g = m.vision_encoder.blocks[1].mlp.fc2.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
