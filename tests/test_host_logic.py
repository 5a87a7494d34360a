"""CPU tests of the host side: C-ABI exports, drop-in module surface (names / state_dict keys / parameter
counts of SURVEY.md App. A.6), flat parameter arena, and the bucketed all-reduce over gloo (world_size 2)."""
import ctypes
import os
import re
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol():
    from x2vlm_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "x2k.h")).read()
    declared = set(re.findall(r"\b(x2k_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.SYMBOLS), (declared ^ set(_capi.SYMBOLS))
    lib = _capi.lib()  # loads libx2k.so and resolves every symbol (AttributeError otherwise)
    assert lib.x2k_version() == 100
    assert ctypes.sizeof(_capi.X2kGemmArgs) > 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "x2vlm_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_ops_refuse_cpu_tensors():
    from x2vlm_b200 import _capi, ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(_capi.X2kError):
        ops.gemm(a, a, 128, 128, 64, out_bf16=torch.zeros(128, 128, dtype=torch.bfloat16))


def _build_xvlm():
    from x2vlm_b200 import pretrain
    torch.manual_seed(0)
    return pretrain.XVLM(pretrain.base_config())


@pytest.fixture(scope="module")
def xvlm():
    return _build_xvlm()


def test_state_dict_surface_matches_reference(xvlm):
    sd = xvlm.state_dict()
    assert len(sd) == 587
    assert sum(p.numel() for p in xvlm.parameters()) == 254758401
    for k, shape in {"temp": (), "vision_encoder.cls_token": (1, 1, 768),
                     "vision_encoder.patch_embed.proj.weight": (768, 3, 16, 16),
                     "vision_encoder.blocks.11.attn.relative_position_bias_table": (732, 12),
                     "vision_encoder.blocks.0.attn.relative_position_index": (197, 197),
                     "vision_encoder.blocks.3.attn.qkv.weight": (2304, 768), "vision_encoder.blocks.3.gamma_1": (768,),
                     "vision_encoder.fc_norm.weight": (768,),
                     "text_encoder.bert.embeddings.position_ids": (1, 512),
                     "text_encoder.bert.embeddings.word_embeddings.weight": (30522, 768),
                     "text_encoder.bert.encoder.layer.0.attention.self.query.weight": (768, 768),
                     "text_encoder.bert.encoder.layer.12.crossattention.self.key.weight": (768, 768),
                     "text_encoder.bert.encoder.layer.17.output.LayerNorm.bias": (768,),
                     "text_encoder.cls.predictions.decoder.weight": (30522, 768),
                     "text_encoder.cls.predictions.decoder.bias": (30522,), "text_encoder.cls.predictions.bias": (30522,),
                     "vision_proj.weight": (256, 768), "itm_head.3.weight": (2, 1536), "bbox_head.0.weight": (1536, 768)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert "text_encoder.bert.encoder.layer.11.crossattention.self.key.weight" not in sd
    # tied decoder
    assert xvlm.text_encoder.cls.predictions.decoder.weight is xvlm.text_encoder.bert.embeddings.word_embeddings.weight
    assert sd["vision_encoder.blocks.0.attn.relative_position_index"].dtype == torch.int64


@pytest.mark.reference
def test_state_dict_keys_identical_to_reference(xvlm):
    from oracle import ref_shim
    ref = ref_shim.build_reference_xvlm()
    rsd, sd = ref.state_dict(), xvlm.state_dict()
    assert list(rsd.keys()) == list(sd.keys()) or set(rsd.keys()) == set(sd.keys())
    for k in rsd:
        assert rsd[k].shape == sd[k].shape and rsd[k].dtype == sd[k].dtype, k
    assert torch.equal(rsd["vision_encoder.blocks.0.attn.relative_position_index"],
                       sd["vision_encoder.blocks.0.attn.relative_position_index"])
    missing, unexpected = xvlm.load_state_dict(rsd, strict=True)
    assert not missing and not unexpected


def test_arena_layout_and_views(xvlm):
    import copy
    from x2vlm_b200.params import ParamArena
    m = copy.deepcopy(xvlm.text_encoder.bert.encoder.layer[12])  # a fusion layer
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    arena = ParamArena(m)
    for n, p in m.named_parameters():
        assert torch.equal(p, before[n]), n
        s, e = arena.span(p)
        assert p.data_ptr() == arena.flat[s:e].data_ptr() and p.grad.data_ptr() == arena.grad[s:e].data_ptr()
        assert s % 8 == 0 or any(p is q for sh in arena.shadows for q in sh.params[1:])
    at = m.attention.self
    q, k, v = (arena.span(x.weight)[0] for x in (at.query, at.key, at.value))
    assert k == q + at.query.weight.numel() and v == k + at.key.weight.numel()  # packed QKV weights are adjacent
    sh = m._x2k["qkv"]
    assert sh.arena is arena and sh.grad_sink().shape == (3 * 768, 768)
    assert sh.grad_sink().data_ptr() == at.query.weight.grad.data_ptr()


def _ddp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from x2vlm_b200.accelerator import GradBucketer
    from x2vlm_b200.params import ParamArena
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.GELU(), torch.nn.Linear(256, 64), torch.nn.LayerNorm(64))
    arena = ParamArena(m)
    b = GradBucketer(arena, world, bucket_mb=0.05)
    assert len(b.buckets) >= 2
    torch.manual_seed(100 + rank)
    x = torch.randn(16, 64)
    arena.zero_grad()
    m(x).pow(2).sum().backward()
    local = arena.grad.clone()
    order = b.finalize()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered) / world
    ok = torch.allclose(arena.grad, want, atol=1e-6) and sorted(order) == list(range(len(b.buckets)))
    # second backward accumulates on top of the averaged gradient (two-backward video pattern, Pretrain.py:193-247)
    m(x).pow(2).sum().backward()
    b.finalize()
    ok = ok and torch.allclose(arena.grad, 2 * want, atol=1e-5)
    q.put((rank, bool(ok), order))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(30) for p in procs]
    assert all(ok for _, ok, _ in res), res
    # buckets complete in reverse parameter order (last layer first): overlap-friendly launch order
    assert res[0][2][0] == max(res[0][2])


def test_flat_adamw_groups(xvlm):
    import copy
    from x2vlm_b200.accelerator import FlatAdamW
    from x2vlm_b200.params import ParamArena
    m = copy.deepcopy(xvlm.text_encoder.bert.encoder.layer[0])
    arena = ParamArena(m)
    opt = FlatAdamW(m, arena, lr=1e-4, weight_decay=0.01)
    wds = sorted({g["weight_decay"] for g in opt.param_groups})
    assert wds == [0.0, 0.01]
    nd = [p for g in opt.param_groups if g["weight_decay"] == 0.0 for p in g["params"]]
    names = {id(p): n for n, p in m.named_parameters()}
    assert all(("bias" in names[id(p)] or "LayerNorm" in names[id(p)]) for p in nd)
    assert int(opt.seg_end[-1]) == arena.numel


def test_retrieval_recall_bookkeeping():
    """itm_eval (Retrieval.py:160-209) on a hand-checked case: ranks of the best ground-truth caption / the image."""
    from x2vlm_b200 import retrieval
    s_i2t = torch.tensor([[0.9, 0.1, 0.3, -100.0], [0.2, 0.8, -100.0, 0.7]])      # 2 images x 4 captions
    s_t2i = torch.tensor([[0.9, 0.1], [0.3, 0.6], [0.8, 0.2], [0.9, 0.1]])        # caption 3 ranks its image second
    r = retrieval.itm_eval(s_i2t, s_t2i, txt2img={0: 0, 1: 1, 2: 0, 3: 1}, img2txt={0: [0, 2], 1: [1, 3]})
    assert r["txt_r1"] == 100.0 and r["img_r1"] == 75.0 and r["img_r5"] == 100.0
    assert abs(r["r_mean"] - (100.0 + (75.0 + 100.0 + 100.0) / 3) / 2) < 1e-9
    assert retrieval._rank_slice(1000) == (0, 1000, 1)   # single process: all rows


def test_generation_surface_on_cpu():
    """BertLMHeadModel constructs with the reference's state_dict keys; label-smoothed CE equals its definition."""
    from x2vlm_b200 import xbert
    cfg = xbert.BertConfig(vocab_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                           max_position_embeddings=32)
    cfg.fusion_layer, cfg.encoder_width = 0, 128
    m = xbert.BertLMHeadModel(cfg, label_smoothing=0.1)
    keys = set(m.state_dict().keys())
    assert "bert.encoder.layer.1.crossattention.self.key.weight" in keys and "cls.predictions.decoder.weight" in keys
    assert m.cls.predictions.decoder.weight is m.bert.embeddings.word_embeddings.weight      # tied
    g = torch.Generator().manual_seed(0)
    logits, labels = torch.randn(7, 64, generator=g), torch.tensor([3, -100, 5, 0, 63, -100, 9])
    got = xbert.LabelSmoothSoftmaxCEV1(0.1, "mean")(logits, labels)
    logp = logits.log_softmax(1)
    t = torch.full_like(logp, 0.1 / 64)
    valid = labels != -100
    t[valid, labels[valid]] = 0.9
    want = -(logp * t).sum(1)[valid].sum() / valid.sum()
    assert abs(float(got) - float(want)) < 1e-6
    out = m.prepare_inputs_for_generation(torch.ones(2, 5, dtype=torch.long), past=((None,),))
    assert out["input_ids"].shape == (2, 1) and out["is_decoder"] is True


def _retrieval_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from x2vlm_b200 import retrieval
    n = 7
    full = torch.arange(n * 5, dtype=torch.float32).view(n, 5) - 10.0
    full[2, 3] = -100.0
    start, end, w = retrieval._rank_slice(n)          # step = 7 // 2 + 1 = 4: rank 0 rows 0-3, rank 1 rows 4-6
    mine = torch.full((n, 5), -100.0)
    mine[start:end] = full[start:end]
    retrieval.combine_rank_rows(mine)
    # the reference's SUM of -100-filled matrices (Retrieval.py:145-148): every entry shifted by -100 * (W - 1)
    q.put((rank, (start, end, w), bool(torch.equal(mine, full - 100.0 * (w - 1)))))
    dist.destroy_process_group()


def test_retrieval_rank_partition_gloo_world2():
    """Rows split over 2 processes like Retrieval.py:120-123 and recombined on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 200
    procs = [ctx.Process(target=_retrieval_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(30) for p in procs]
    assert res[0][1] == (0, 4, 2) and res[1][1] == (4, 7, 2)
    assert res[0][2] and res[1][2]


def test_beit2_checkpoint_helpers(tmp_path):
    """load_pretrained_beit2 / interpolate_pos_embed / load_state_dict (models/beit2.py:473-754) on CPU."""
    from functools import partial
    from x2vlm_b200 import beit2
    mk = lambda res: beit2.VisionTransformer(img_size=res, patch_size=16, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4,
                                             norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), use_rel_pos_bias=True,
                                             use_abs_pos_emb=False, init_values=0.1, qkv_bias=True)
    src = mk(224)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in src.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    # BEiT-2 release layout: wrapped in 'model', classification head, ONE shared rel-pos table
    ckpt = {k: v for k, v in sd.items() if "relative_position_bias_table" not in k}
    ckpt["rel_pos_bias.relative_position_bias_table"] = sd["blocks.0.attn.relative_position_bias_table"].clone()
    ckpt["head.weight"], ckpt["head.bias"] = torch.zeros(10, 128), torch.zeros(10)
    path = tmp_path / "beit2.pth"
    torch.save({"model": ckpt}, path)
    # same resolution: every tensor arrives unchanged, the index buffer is rebuilt by the model, the head is dropped
    dst = mk(224)
    missing, unexpected, ignored = beit2.load_pretrained_beit2(dst, str(path))
    assert missing == [] and unexpected == [] and all("relative_position_index" in k for k in ignored)
    for k, v in dst.state_dict().items():
        want = sd["blocks.0.attn.relative_position_bias_table"] if "relative_position_bias_table" in k else sd[k]
        assert torch.equal(v, want), k
    # 224 -> 384: 27x27 (+3) tables become 47x47 (+3)
    big = mk(384)
    sd2 = beit2.interpolate_pos_embed(big, {k: v.clone() for k, v in sd.items()})
    assert not any("relative_position_index" in k for k in sd2)
    t_src, t_dst = sd["blocks.1.attn.relative_position_bias_table"], sd2["blocks.1.attn.relative_position_bias_table"]
    assert t_src.shape == (27 * 27 + 3, 2) and t_dst.shape == (47 * 47 + 3, 2)
    assert torch.equal(t_dst[-3:], t_src[-3:])                       # cls rows are copied
    # an interpolating spline reproduces the source where target offsets coincide with source offsets: 0 and +-1
    s_img, d_img = t_src[:-3, 0].view(27, 27), t_dst[:-3, 0].view(47, 47)
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            assert abs(float(d_img[23 + a, 23 + b]) - float(s_img[13 + a, 13 + b])) < 1e-4
    assert float(d_img.abs().max()) < 3 * float(s_img.abs().max())   # no wild overshoot
    missing, unexpected, _ = beit2.load_state_dict(big, sd2)
    assert missing == [] and unexpected == []
    # absolute position embedding: bicubic resize of the patch grid, cls token kept
    pe = mk(224)
    pe.pos_embed = torch.nn.Parameter(torch.zeros(1, 197, 128))
    pe_big = mk(384)
    pe_big.pos_embed = torch.nn.Parameter(torch.zeros(1, 577, 128))
    out = beit2.interpolate_pos_embed(pe_big, {"pos_embed": torch.randn(1, 197, 128, generator=g)})
    assert out["pos_embed"].shape == (1, 577, 128)


def test_generate_no_beam_host_loop():
    """BertLMHeadModel.generate_no_beam (models/xbert.py:1415-1519) with the network replaced by a scripted logits
    table: greedy path, EOS bookkeeping, padding, repetition penalty and the cache hand-over are host logic."""
    from types import SimpleNamespace
    from x2vlm_b200 import xbert
    cfg = xbert.BertConfig(vocab_size=12, hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256,
                           max_position_embeddings=32)
    cfg.fusion_layer, cfg.encoder_width = 0, 128
    m = xbert.BertLMHeadModel(cfg)
    calls = []

    def fake_forward(input_ids=None, past_key_values=None, use_cache=None, return_dict=None, **kw):
        calls.append((tuple(input_ids.shape), past_key_values))
        step = len(calls)
        logits = torch.full((input_ids.shape[0], input_ids.shape[1], 12), -5.0)
        # sequence 0 emits 3, 4, then EOS (= 2); sequence 1 emits 5 forever
        plan0 = {1: 3, 2: 4}.get(step, 2)
        logits[0, -1, plan0] = 5.0
        logits[1, -1, 5] = 5.0
        logits[1, -1, 6] = 4.9          # runner-up: wins once 5 is penalised
        return SimpleNamespace(logits=logits, past_key_values=("cache", step))

    m.forward = fake_forward
    ids, lp = m.generate_no_beam(torch.tensor([[1, 7], [1, 8]]), max_length=7, eos_token_ids=(2,), pad_token_id=0)
    assert ids.tolist() == [[1, 7, 3, 4, 2, 0, 0], [1, 8, 5, 5, 5, 5, 2]]        # finished rows are padded; EOS forced at the end
    assert calls[0] == ((2, 2), None) and calls[1] == ((2, 1), ("cache", 1))    # prompt once, then one token per step
    assert lp.shape == (2,) and torch.isfinite(lp).all()
    calls.clear()
    ids, _ = m.generate_no_beam(torch.tensor([[1, 7], [1, 8]]), max_length=5, eos_token_ids=(2,), repetition_penalty=1.5)
    assert ids[1].tolist()[2:4] == [5, 6]                                         # 5 / 1.5 < 4.9 after its first use


def test_bench_reference_arm_prints_one_json_line():
    """Contract: `bench.py --impl reference` prints exactly ONE JSON line on stdout (library chatter goes to stderr),
    with the same metric / unit / workload as the main arm and a cpu_baseline describing the run."""
    import json
    import subprocess
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "image-text pairs/sec pretrain step X2VLM-base bf16"
    assert d["config"]["workload"].startswith("X2VLM-base pretrain step (ITC+ITM+MLM+bbox), 224px, 40 tok, batch 64/GPU")
    # the unmodified reference itself (oracle/_ref byte-code, built by __graft_entry__.build()), not the port
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert "UNMODIFIED reference" in d["cpu_baseline"]["sample"]
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    assert d["config"] == bench.workload_config(argparse.Namespace(batch=64, region_images=26, image_only=False), 1)
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.reference
def test_set_up_takes_over_reference_optimizer_and_scheduler(xvlm):
    """The drop-in flow of Pretrain.py:560-579: optim.create_optimizer -> scheduler.create_scheduler -> accelerator.set_up
    -> reinit_scheduler_properties_mysched.  The flat optimizer takes over the reference AdamW's groups one to one, is a
    real torch Optimizer (LambdaLR accepts it), follows the scheduler, and round-trips its state."""
    import copy
    import types
    from oracle import ref_shim
    from x2vlm_b200.accelerator import FlatAdamW, rebind_scheduler
    from x2vlm_b200.params import ParamArena
    ref_shim.install()
    import optim as ref_optim
    import scheduler as ref_sched

    class AttrDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    m = torch.nn.Module()
    m.vision_encoder = copy.deepcopy(xvlm.vision_encoder.blocks[0])
    m.text_encoder = copy.deepcopy(xvlm.text_encoder.bert.encoder.layer[12])
    m.head = torch.nn.Linear(8, 8)
    m.text_encoder.intermediate.dense.weight.requires_grad_(False)      # a frozen weight (fine-tuning)
    args = types.SimpleNamespace(lr=1e-4, weight_decay=0.01, lr_mult=2, vision_lr=2e-5, text_lr=4e-5)
    opt = ref_optim.create_optimizer(args, m)                           # the reference's own grouping (optim.py:26-104)
    sch = ref_sched.create_scheduler(AttrDict(sched="linear", epochs=1, step_per_epoch=100, num_warmup_steps=10), opt)
    arena = ParamArena(m)
    flat = FlatAdamW.from_optimizer(m, arena, opt)
    assert isinstance(flat, torch.optim.Optimizer) and len(flat.param_groups) == len(opt.param_groups) == 10
    for g0, g1 in zip(opt.param_groups, flat.param_groups):
        assert [id(p) for p in g0["params"]] == [id(p) for p in g1["params"]]
        assert g1["weight_decay"] == g0["weight_decay"] and tuple(g1["betas"]) == (0.9, 0.98) and g1["eps"] == 1e-8
    sch2 = rebind_scheduler(sch, flat)
    assert sch2.optimizer is flat
    # Pretrain.py:33-51 re-initialises the scheduler on the optimizer set_up returned
    assert sch2.optimizer == flat
    sch2.__init__(flat, sch2.lr_lambdas[0], last_epoch=-1)
    assert all(g["lr"] == 0.0 for g in flat.param_groups)               # warm-up starts at 0
    for _ in range(5):
        sch2.step()
    lrs = [g["lr"] for g in flat.param_groups]
    assert abs(lrs[0] - 0.5e-4) < 1e-12 and abs(lrs[4] - 1e-5) < 1e-12 and abs(lrs[6] - 2e-5) < 1e-12
    flat._upload_hparams()
    names = {id(p): n for n, p in m.named_parameters()}
    for i, p in enumerate(arena.params):
        want = 0.0 if not p.requires_grad else (1e-5 if names[id(p)].startswith("vision_encoder") else
                                                2e-5 if names[id(p)].startswith("text_encoder") else 0.5e-4)
        assert abs(float(flat.seg_lr[i]) - want) < 1e-10, names[id(p)]
    # checkpoint round trip (training_states, Pretrain.py:603)
    flat.exp_avg.normal_(); flat.exp_avg_sq.uniform_(); flat.step_dev.fill_(7)
    sd = flat.state_dict()
    other = FlatAdamW.from_optimizer(m, arena, ref_optim.create_optimizer(args, m))
    other.load_state_dict(sd)
    assert torch.equal(other.exp_avg, flat.exp_avg) and int(other.step_dev) == 7
    assert [g["lr"] for g in other.param_groups] == lrs
    # anything that is not AdamW is rejected loudly
    with pytest.raises(TypeError):
        FlatAdamW.from_optimizer(m, arena, torch.optim.SGD(m.parameters(), lr=0.1))


def test_arena_gapped_beit_bias(xvlm):
    """BEiT's QKV bias [q_bias | 0 | v_bias] (models/beit2.py:129) is a plain view of the arena: the two parameters are laid
    out with a zero hole of one bias length between them, gradients land in their own slices."""
    import copy
    from x2vlm_b200.params import GappedBias, ParamArena
    blk = copy.deepcopy(xvlm.vision_encoder.blocks[2])
    with torch.no_grad():
        blk.attn.q_bias.copy_(torch.arange(768.0))
        blk.attn.v_bias.copy_(-torch.arange(768.0))
    arena = ParamArena(blk)
    g = blk._x2k["bqv"]
    assert isinstance(g, GappedBias) and g.arena is arena
    packed = g.get_f32()
    assert packed.shape == (3 * 768,)
    assert torch.equal(packed[:768], blk.attn.q_bias) and torch.equal(packed[1536:], blk.attn.v_bias)
    assert float(packed[768:1536].abs().sum()) == 0.0
    gq, gv = g.grad_views()
    assert gq.data_ptr() == blk.attn.q_bias.grad.data_ptr() and gv.data_ptr() == blk.attn.v_bias.grad.data_ptr()
    sq, eq = arena.span(blk.attn.q_bias)
    sv, ev = arena.span(blk.attn.v_bias)
    assert sv == eq + 768 and sq % 8 == 0


def test_gather_rows_backward_matches_slicing():
    """pretrain._gather_rows (one gather of every row the loss heads read from the fusion output): same values and the
    same input gradient as the slices / torch.gather it replaced, including repeated rows (padded masked positions)."""
    from x2vlm_b200 import pretrain
    torch.manual_seed(0)
    S, L, D = 6, 5, 4
    x = torch.randn(S, L, D, requires_grad=True)
    pos = torch.tensor([[1, 3, 0], [2, 2, 4]])              # masked positions of sequences 4 and 5 (a repeat on purpose)
    cls_seq = torch.tensor([0, 1, 2, 3])
    idx = torch.cat([cls_seq * L, ((4 + torch.arange(2)) * L).unsqueeze(1).add(pos).reshape(-1)])
    rows = pretrain._gather_rows(x, idx)
    w = torch.randn_like(rows)
    (rows * w).sum().backward()
    got = x.grad.clone()
    x.grad = None
    ref_rows = torch.cat([x[:4, 0], torch.gather(x[4:], 1, pos.unsqueeze(2).expand(-1, -1, D)).reshape(-1, D)])
    assert torch.equal(rows, ref_rows)
    (ref_rows * w).sum().backward()
    assert torch.allclose(got, x.grad, atol=1e-6)
