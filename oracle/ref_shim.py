"""TEST INFRASTRUCTURE — import shim that lets the UNMODIFIED reference (/root/reference) run on CPU.

Only tests/, oracle/make_golden.py and the validation of oracle/restate.py use this; nothing in the
product path (x2vlm_b200/) may import anything under oracle/.  The reference tree only exists in the
build container (not on the GPU box), so everything here is guarded by `available()`.

What it bridges: the reference pins timm 0.4.9 and transformers 4.12.5
and imports ruamel.yaml / pycocotools / pycocoevalcap / skimage / matplotlib at module import time;
none of these are installed (transformers is 5.x).  We insert small stand-in modules and alias the
handful of moved transformers helpers — the reference's own files are not touched.
"""
import json
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("X2VLM_REFERENCE_ROOT", "/root/reference")

_installed = False
_workdir = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _fake_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Permissive(types.ModuleType):
    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return type(item, (), {})


def install():
    """Insert the stand-in modules and transformers aliases (idempotent)."""
    global _installed
    if _installed:
        return
    import torch
    import transformers  # noqa: F401  (must be imported before the fake timm goes in)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    import transformers.optimization
    import transformers.file_utils as fu
    import yaml

    # ---- timm 0.4.9 stand-ins (drop_path: per-sample mask, x.div(keep) * floor(keep + rand)) ----
    def drop_path(x, drop_prob: float = 0.0, training: bool = False):
        if drop_prob == 0.0 or not training:
            return x
        keep_prob = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        random_tensor = keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device)
        random_tensor.floor_()
        return x.div(keep_prob) * random_tensor

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    timm = _fake_module("timm")
    timm_models = _fake_module("timm.models")
    timm_layers = _fake_module("timm.models.layers", drop_path=drop_path, to_2tuple=to_2tuple,
                               trunc_normal_=torch.nn.init.trunc_normal_)

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=None):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return drop_path(x, self.drop_prob, self.training)

    timm_layers.DropPath = DropPath
    timm_registry = _fake_module("timm.models.registry", register_model=lambda f: f)
    timm.models = timm_models
    timm_models.layers = timm_layers
    timm_models.registry = timm_registry

    # ---- misc import-time dependencies ----
    ruamel = _fake_module("ruamel")
    ruamel.yaml = _fake_module("ruamel.yaml", load=yaml.load, dump=yaml.dump, Loader=yaml.Loader)
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.mask", "pycocoevalcap", "pycocoevalcap.eval",
                 "skimage", "skimage.io", "matplotlib", "matplotlib.pyplot", "matplotlib.collections",
                 "matplotlib.patches", "cv2", "hdfs"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Permissive(name)
    for parent, child in (("pycocotools", "coco"), ("pycocotools", "mask"), ("pycocoevalcap", "eval"), ("skimage", "io"),
                          ("matplotlib", "pyplot"), ("matplotlib", "collections"), ("matplotlib", "patches")):
        p = sys.modules.get(parent)
        if isinstance(p, _Permissive):
            setattr(p, child, sys.modules[parent + "." + child])

    # ---- transformers 4.12.5 names on a 5.x install ----
    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer

    def _no_prune(*a, **k):
        raise NotImplementedError("head pruning is not part of the hot path")

    mu.find_pruneable_heads_and_indices = _no_prune
    mu.get_parameter_dtype = lambda m: next(m.parameters()).dtype
    transformers.optimization.AdamW = torch.optim.AdamW
    ident = lambda *a, **k: (lambda f: f)
    for n in ("add_code_sample_docstrings", "replace_return_docstrings", "add_start_docstrings",
              "add_start_docstrings_to_model_forward"):
        setattr(fu, n, ident)
    mu.PreTrainedModel.get_head_mask = lambda self, hm, n, *a: [None] * n

    def init_weights(self):  # 4.12.5 semantics: apply _init_weights, then tie output <- input embeddings
        self.apply(self._init_weights)
        out = self.get_output_embeddings() if hasattr(self, "get_output_embeddings") else None
        if out is not None and getattr(self.config, "tie_word_embeddings", True):
            out.weight = self.get_input_embeddings().weight

    mu.PreTrainedModel.init_weights = init_weights

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


BERT_BASE_CONFIG = {
    "architectures": ["BertForMaskedLM"], "attention_probs_dropout_prob": 0.1, "hidden_act": "gelu",
    "hidden_dropout_prob": 0.1, "hidden_size": 768, "initializer_range": 0.02, "intermediate_size": 3072,
    "layer_norm_eps": 1e-12, "max_position_embeddings": 512, "model_type": "bert", "num_attention_heads": 12,
    "num_hidden_layers": 12, "pad_token_id": 0, "type_vocab_size": 2, "vocab_size": 30522,
}


def workdir():
    """Scratch cwd holding configs/config_beit2_*.json (copied) and a synthesised data/bert-base-uncased."""
    global _workdir
    if _workdir is None:
        d = tempfile.mkdtemp(prefix="x2vlm_oracle_")
        os.makedirs(os.path.join(d, "configs"))
        for n in ("config_beit2_base.json", "config_beit2_large.json"):
            src = os.path.join(REFERENCE_ROOT, "configs", n)
            if os.path.exists(src):
                with open(src) as fi, open(os.path.join(d, "configs", n), "w") as fo:
                    fo.write(fi.read())
        bd = os.path.join(d, "data", "bert-base-uncased")
        os.makedirs(bd)
        with open(os.path.join(bd, "config.json"), "w") as f:
            json.dump(BERT_BASE_CONFIG, f)
        vocab = ["[unused%d]" % i for i in range(30522)]
        vocab[0], vocab[100], vocab[101], vocab[102], vocab[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
        with open(os.path.join(bd, "vocab.txt"), "w") as f:
            f.write("\n".join(vocab) + "\n")
        _workdir = d
    return _workdir


def base_config(**over):
    cfg = dict(use_beit_v2=True, vision_config="configs/config_beit2_base.json", image_res=224, patch_size=16,
               local_attn_depth=-1, text_encoder="data/bert-base-uncased", text_num_hidden_layers=18,
               text_fusion_start_at=12, embed_dim=256, temp=0.07, accelerator={"FP16_OPT_LEVEL": "O0"})
    cfg.update(over)
    return cfg


def init_dist():
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("gloo", rank=0, world_size=1)


def build_reference_xvlm(config=None, seed=0):
    """Construct the reference models.model_pretrain.XVLM (random init, seeded) on CPU."""
    import torch
    install()
    init_dist()
    cwd = os.getcwd()
    os.chdir(workdir())
    try:
        from models.model_pretrain import XVLM
        torch.manual_seed(seed)
        m = XVLM(config or base_config(), load_vision_params=False, load_text_params=False, pretraining=False)
    finally:
        os.chdir(cwd)
    return m
