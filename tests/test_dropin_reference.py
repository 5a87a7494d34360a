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


SCRIPT_HEADS = r'''
import os, sys, tempfile
sys.path.insert(0, %(root)r)
from oracle import ref_shim
ref_shim.install(); ref_shim.init_dist()
import torch
from x2vlm_b200 import beit2, xbert
sys.modules["models.beit2"] = beit2
sys.modules["models.xbert"] = xbert
os.chdir(ref_shim.workdir())
from models.model_pretrain import XVLM
from models.model_retrieval import XVLMForRetrieval
from models.model_generation import XVLMForVQA, XVLMForMLMCaptioning
cfg = ref_shim.base_config()
torch.manual_seed(0)
pre = XVLM(cfg, load_vision_params=False, load_text_params=False, pretraining=False)
ck = os.path.join(tempfile.mkdtemp(), "pre.th")
torch.save({"model": pre.state_dict()}, ck)
# retrieval head: BertModel text encoder; the reference's load_pretrained renames 'text_encoder.bert.*' and adapts the
# vision tables through models.beit2.interpolate_pos_embed -- here x2vlm_b200.beit2's
ret = XVLMForRetrieval(cfg)
assert type(ret.text_encoder).__module__ == "x2vlm_b200.xbert" and type(ret.text_encoder).__name__ == "BertModel"
ret.load_pretrained(ck, cfg, is_eval=False)
a, b = pre.state_dict(), ret.state_dict()
for k, v in b.items():
    src = k if k in a else k.replace("text_encoder.", "text_encoder.bert.", 1)
    assert src in a and torch.equal(a[src], v), k
assert ret.init_params == []
# VQA: encoder + BertLMHeadModel answer decoder (fusion_layer 0, 6 layers); captioning: BertForMaskedLM + prompt
vcfg = dict(cfg, pad_token_id=0, num_dec_layers=6)
vqa = XVLMForVQA(vcfg)
assert type(vqa.text_decoder).__module__ == "x2vlm_b200.xbert" and type(vqa.text_decoder).__name__ == "BertLMHeadModel"
assert vqa.text_decoder.config.fusion_layer == 0 and len(vqa.text_decoder.bert.encoder.layer) == 6
assert all(l.has_cross_attention for l in vqa.text_decoder.bert.encoder.layer)
cap = XVLMForMLMCaptioning(dict(cfg, prompt="a picture of", label_smoothing=0.1))
assert type(cap.text_encoder).__name__ == "BertForMaskedLM" and len(cap.prompt_ids) >= 2
# grounding (bbox head; text / cross DropPath as in configs/finetune/refcoco_grounding_large.yaml) and classification heads
from models.model_grounding import XVLMForGrounding
from models.model_classification import XVLMForClassification, XVLMForVQAClassification
gcfg = dict(cfg, text_drop_path_rate=0.1, cross_drop_path_rate=0.1)
grd = XVLMForGrounding(gcfg)
assert isinstance(grd.text_encoder.encoder.layer[17].output.drop_path, xbert.DropPath)
assert grd.text_encoder.config.hidden_dropout_prob == 0.0 and hasattr(grd, "bbox_head")
grd.load_pretrained(ck, gcfg)
assert torch.equal(grd.bbox_head[0].weight, pre.bbox_head[0].weight)
assert torch.equal(grd.text_encoder.encoder.layer[3].output.dense.weight, pre.text_encoder.bert.encoder.layer[3].output.dense.weight)
cls = XVLMForClassification(dict(cfg, num_labels=3))
vqc = XVLMForVQAClassification(dict(cfg, num_labels=10))
assert cls.cls_head[-1].out_features == 3 and len(vqc.init_params) > 0
print("HEADS_OK")
'''


@pytest.mark.reference
def test_reference_task_heads_construct_and_load_on_x2k_encoders():
    """models/model_retrieval.py and models/model_generation.py (unmodified) on the B200-native modules: construction,
    and the reference's own checkpoint flow (pretrain checkpoint -> XVLMForRetrieval.load_pretrained)."""
    res = subprocess.run([sys.executable, "-c", SCRIPT_HEADS % {"root": ROOT}], capture_output=True, text=True, timeout=900)
    assert "HEADS_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
