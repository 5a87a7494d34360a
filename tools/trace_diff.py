"""Which kernel output first differs between the first and the second identical pass?  Every x2vlm_b200.ops call is
wrapped; after it returns all tensor arguments are hashed (sum of fp64 + sum of squares); pass 0 and pass 1 traces are
compared in order.  (Developer tool: localises kernels that read uninitialised memory.)"""
import os, sys, inspect
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import accelerator, ops, pretrain, synth
dev = torch.device("cuda:0")
trace = []
NAMES = ["gemm", "layernorm_fwd", "layernorm_bwd", "scale_cast_colsum", "colsum_bf16", "segment_sum_bf16", "attn_fwd", "attn_bwd",
         "relpos_bias_gather", "relpos_bias_scatter", "cast_f32_bf16"]
def sig(t):
    t = t.detach().double()
    return (float(t.sum()), float((t * t).sum()))
def wrap(name, fn):
    params = list(inspect.signature(fn).parameters)
    def inner(*a, **k):
        r = fn(*a, **k)
        rec = {}
        for i, v in enumerate(a):
            if isinstance(v, torch.Tensor): rec[params[i] if i < len(params) else "arg%d" % i] = sig(v)
            elif isinstance(v, (tuple, list)):
                for j, w in enumerate(v):
                    if isinstance(w, torch.Tensor): rec["%s[%d]" % (params[i], j)] = sig(w)
        for kk, v in k.items():
            if isinstance(v, torch.Tensor): rec[kk] = sig(v)
        shp = [x for x in a if isinstance(x, int)][:3]
        if name == "cast_f32_bf16" and a[0].numel() > (1 << 26):
            return r
        trace.append((name, shp, rec))
        return r
    return inner
for n in NAMES:
    setattr(ops, n, wrap(n, getattr(ops, n)))
torch.manual_seed(0)
m = pretrain.XVLM(pretrain.base_config(vision_num_hidden_layers=2, text_num_hidden_layers=4, text_fusion_start_at=2))
acc = accelerator.X2kDDPAccelerator({"lr": 1e-4, "weight_decay": 0.01})
ddp, opt, _ = acc.set_up(m, None, None, 0, 1, 0)
ddp.eval()
acc.arena.sync_bf16()  # keep the op lists of pass 0 and pass 1 aligned
B = 4
ib = {k: v.to(dev) for k, v in synth.image_text_batch(B, 40, seed=100).items()}
neg = tuple(t.to(dev) for t in synth.hard_negative_indices(B, 5))
def run():
    trace.clear()
    opt.zero_grad()
    loss = ddp.module.total_loss(ddp.module.forward_mixed(ib, None, neg))
    acc.backward_step(loss, opt)
    torch.cuda.synchronize()
    return list(trace), acc.arena.grad.clone()
t0, g0 = run()
t1, g1 = run()
t2, g2 = run()
print("grad diff 0-1 %.3e  1-2 %.3e   ops %d %d" % (((g0 - g1).norm() / g0.norm()).item(), ((g1 - g2).norm() / g1.norm()).item(), len(t0), len(t1)))
shown = 0
for i, (a, b) in enumerate(zip(t0, t1)):
    bad = [k for k in a[2] if k in b[2] and (abs(a[2][k][0] - b[2][k][0]) > 1e-9 * (abs(a[2][k][0]) + 1e-30) + 0 or abs(a[2][k][1] - b[2][k][1]) > 1e-9 * abs(a[2][k][1]))]
    if bad:
        print("op %d %s %s differs in %s" % (i, a[0], a[1], [(k, a[2][k], b[2][k]) for k in bad][:3]))
        shown += 1
        if shown >= 6:
            break
