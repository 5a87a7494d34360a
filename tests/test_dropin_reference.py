"""Drop-in boundary check (build container only): the reference's UNMODIFIED models/xvlm.py + models/model_pretrain.py
construct their XVLM on top of x2vlm_b200.beit2 / x2vlm_b200.xbert when those are pre-seeded as `models.beit2` /
`models.xbert` (SURVEY.md §7.1 patch point (i)).  Runs in a fresh interpreter because other tests import the real
reference modules."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
from oracle import ref_shim
ref_shim.install(); ref_shim.init_dist()
import torch
from x2vlm_b200 import beit2, xbert, pretrain
sys.modules["models.beit2"] = beit2
sys.modules["models.xbert"] = xbert
os.chdir(ref_shim.workdir())
from models.model_pretrain import XVLM          # the reference's own class
import models.xvlm as rx
assert rx.BertForMaskedLM is xbert.BertForMaskedLM and rx.BertModel is xbert.BertModel
torch.manual_seed(0)
m = XVLM(ref_shim.base_config(), load_vision_params=False, load_text_params=False, pretraining=False)
assert type(m.vision_encoder).__module__ == "x2vlm_b200.beit2", type(m.vision_encoder)
assert type(m.text_encoder).__module__ == "x2vlm_b200.xbert", type(m.text_encoder)
assert sum(p.numel() for p in m.parameters()) == 254758401
mine = pretrain.XVLM(pretrain.base_config())
a, b = m.state_dict(), mine.state_dict()
assert set(a) == set(b) and len(a) == 587
assert all(a[k].shape == b[k].shape for k in a)
assert m.vision_encoder.vision_width == 768 and m.text_encoder.config.fusion_layer == 12
# the reference's caller-side methods resolve the attributes they read
assert m.text_encoder.bert.embeddings.word_embeddings.weight is m.text_encoder.cls.predictions.decoder.weight
assert hasattr(m.vision_encoder.patch_embed, "num_patches") and m.vision_encoder.patch_embed.num_patches == 196
print("DROPIN_OK")
'''


@pytest.mark.reference
def test_reference_xvlm_constructs_on_x2k_encoders():
    res = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert "DROPIN_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
