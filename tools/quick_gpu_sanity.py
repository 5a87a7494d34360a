"""Seconds-long GPU sanity of the caller-side mirror on a 2+2+1-layer model: mixed step with device-side hard-negative
sampling, retrieval fine-tuning step with sample ids, backward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
import torch
from x2vlm_b200 import pretrain, synth
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = pretrain.XVLM(pretrain.base_config(vision_num_hidden_layers=2, text_num_hidden_layers=3, text_fusion_start_at=2)).to(dev).train()
ib = {k: v.to(dev) for k, v in synth.image_text_batch(6, 40, seed=1).items()}
rb = {k: v.to(dev) for k, v in synth.region_batch(3, 6, 40, seed=2).items()}
loss = m.total_loss(m.forward_mixed(ib, rb))
loss.backward()
li, lm = m.forward_retrieval(ib["image"], ib["text_ids"], ib["text_atts"], idx=torch.tensor([0, 1, 0, 2, 1, 3], device=dev))
(li + lm).backward()
torch.cuda.synchronize()
print("QUICK_OK mixed %.4f itc %.4f itm %.4f  (%.1f s)" % (float(loss.detach()), float(li.detach()), float(lm.detach()), time.time() - t0))
