"""GPU test of the accelerator's drop-in surface (accelerators/apex_ddp_accelerator.py:42-102 as called from
Pretrain.py:560-579, :67-72): set_up with a torch AdamW + LambdaLR built the reference's way, backward_step,
optimizer_step (clip), optimizer.step, scheduler.step — on one GPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_set_up_with_torch_adamw_and_lambdalr():
    from torch.optim.lr_scheduler import LambdaLR
    from x2vlm_b200 import accelerator, pretrain, synth
    torch.manual_seed(0)
    cfg = pretrain.base_config(vision_num_hidden_layers=1, text_num_hidden_layers=2, text_fusion_start_at=1)
    m = pretrain.XVLM(cfg)
    frozen = m.text_encoder.bert.encoder.layer[0].intermediate.dense.weight
    frozen.requires_grad_(False)
    decay = [p for n, p in m.named_parameters() if p.requires_grad and "bias" not in n and "norm" not in n.lower()]
    no_decay = [p for n, p in m.named_parameters() if p.requires_grad and ("bias" in n or "norm" in n.lower())]
    opt0 = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01, "lr": 1e-3}, {"params": [], "weight_decay": 0.0, "lr": 1e-3},
                              {"params": no_decay, "weight_decay": 0.0, "lr": 2e-3}], lr=1e-3, eps=1e-8, betas=(0.9, 0.98))
    sch0 = LambdaLR(opt0, lambda step: min(1.0, step / 2.0), last_epoch=-1)          # lr 0 at step 0 (linear warm-up)
    acc = accelerator.X2kDDPAccelerator({})
    ddp, opt, sch = acc.set_up(m, opt0, sch0, 0, 1, 0)
    assert isinstance(opt, accelerator.FlatAdamW) and isinstance(opt, torch.optim.Optimizer) and sch.optimizer is opt
    assert hasattr(ddp, "module") and len(opt.param_groups) == 3
    ddp.eval()
    dev = torch.device("cuda:0")
    ib = {k: v.to(dev) for k, v in synth.image_text_batch(4, 40, seed=3).items()}
    neg = tuple(t.to(dev) for t in synth.hard_negative_indices(4, 5))

    def step():
        opt.zero_grad()
        loss = ddp.module.total_loss(ddp.module.forward_mixed(ib, None, neg))
        acc.backward_step(loss, opt)
        norm = acc.optimizer_step(opt, ddp, 1.0)
        opt.step()
        sch.step()
        return float(loss.detach()), float(norm)

    before = acc.arena.flat.clone()
    l0, n0 = step()                                     # lr == 0: AdamW leaves every parameter where it was
    assert torch.equal(acc.arena.flat, before) and n0 > 0
    s0, e0 = acc.arena.span(frozen)
    assert float(acc.arena.grad[s0:e0].abs().sum()) == 0.0   # frozen weight: no wgrad, nothing in the clip norm
    l1, _ = step()                                      # lr == 0.5 * base now
    moved = (acc.arena.flat - before).abs()
    assert float(moved.max()) > 0 and float(moved[s0:e0].max()) == 0.0
    assert float(opt.exp_avg[s0:e0].abs().sum()) == 0.0
    # group learning rates reached the kernel: decay group 0.5e-3, no-decay group 1e-3 (Adam's first real step moves by ~lr)
    wq = ddp.module.text_encoder.bert.encoder.layer[1].attention.self.query
    sw, ew = acc.arena.span(wq.weight)
    sb, eb = acc.arena.span(wq.bias)
    assert 0.3e-3 < float(moved[sw:ew].max()) < 0.8e-3 and 0.6e-3 < float(moved[sb:eb].max()) < 1.6e-3
    l2, n2 = step()
    assert l2 == l2 and n2 == n2 and abs(l2) < 1e4      # finite after a full-lr step
    # resume: state round trip into a fresh optimizer
    sd = opt.state_dict()
    opt.exp_avg.zero_()
    opt.load_state_dict(sd)
    assert float(opt.exp_avg.abs().sum()) > 0 and int(opt.step_dev) == sd["step"] == 3
