"""Two identical forward+backward passes (eval mode: no dropout) of the mixed step: which gradients differ run to run?
Expected: only reordering noise of fp32 atomics (split-K wgrad, column sums) ~1e-6; anything larger localises a
kernel whose intermediate depends on scheduling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import accelerator, pretrain, synth
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pretrain.XVLM(pretrain.base_config())
acc = accelerator.X2kDDPAccelerator({"lr": 1e-4, "weight_decay": 0.01})
ddp, opt, _ = acc.set_up(m, None, None, 0, 1, 0)
ddp.eval()
B = int(os.environ.get("B", "4"))
ib = {k: v.to(dev) for k, v in synth.image_text_batch(B, 40, seed=100).items()}
rb = {k: v.to(dev) for k, v in synth.region_batch(max(1, B // 2), B, 40, seed=101).items()} if os.environ.get("REGION") else None
neg = tuple(t.to(dev) for t in synth.hard_negative_indices(B, 5))
negr = tuple(t.to(dev) for t in synth.hard_negative_indices(B, 6)) if rb is not None else None
def run():
    opt.zero_grad()
    loss = ddp.module.total_loss(ddp.module.forward_mixed(ib, rb, neg, negr))
    acc.backward_step(loss, opt)
    torch.cuda.synchronize()
    return acc.arena.grad.clone(), float(loss)
g0, l0 = run()
for it in range(3):
    g1, l1 = run()
    print("pass %d: loss %.6f vs %.6f   grad rel-L2 diff %.3e" % (it + 1, l1, l0, ((g1 - g0).norm() / g0.norm()).item()))
worst = []
for n, p in m.named_parameters():
    s, e = acc.arena.span(p)
    a, b = g0[s:e], g1[s:e]
    d = ((a - b).norm() / a.norm().clamp_min(1e-20)).item()
    worst.append((d, n))
worst.sort(reverse=True)
for d, n in worst[:25]:
    print("%.3e  %s" % (d, n))
