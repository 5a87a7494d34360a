"""torch.profiler view of ONE eager mixed step: which ATen ops (outside the x2k kernels) cost GPU time, by input shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from x2vlm_b200 import accelerator, pretrain, synth
from x2vlm_b200 import functional as XF
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pretrain.XVLM(pretrain.base_config())
acc = accelerator.X2kDDPAccelerator({"lr": 1e-4, "weight_decay": 0.01})
ddp, opt, _ = acc.set_up(m, None, None, 0, 1, 0)
ddp.train()
XF.manual_seed(1)
ib = {k: v.to(dev) for k, v in synth.image_text_batch(64, 40, seed=1234).items()}
rb = {k: v.to(dev) for k, v in synth.region_batch(26, 64, 40, seed=4321).items()}
def step():
    opt.zero_grad()
    loss = ddp.module.total_loss(ddp.module.forward_mixed(ib, rb))
    acc.backward_step(loss, opt)
    acc.optimizer_step(opt, ddp, 1.0)
    opt.step()
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    st = getattr(e, "self_device_time_total", None)
    if st is None:
        st = getattr(e, "self_cuda_time_total", 0)
    if st > 0:
        rows.append((st, e.count, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("self device time total %.2f ms over %d (op, shape) groups" % (tot / 1e3, len(rows)))
for st, n, k, shp in rows[:70]:
    print("%8.1f us  x%-4d %-42s %s" % (st, n, k[:42], shp))
