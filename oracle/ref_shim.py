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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """The reference's sources (build container) or, where they do not exist (GPU box), the byte-code-only copy that
    oracle/build_ref.py compiled from them into oracle/_ref/."""
    env = os.environ.get("X2VLM_REFERENCE_ROOT")
    for cand in (env, "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and (os.path.exists(os.path.join(cand, "models", "__init__.py")) or
                     os.path.exists(os.path.join(cand, "models", "__init__.pyc"))):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()
COMPILED_ONLY = not os.path.exists(os.path.join(REFERENCE_ROOT, "models", "__init__.py"))

_installed = False
_workdir = None


def available():
    return (os.path.exists(os.path.join(REFERENCE_ROOT, "models", "__init__.py")) or
            os.path.exists(os.path.join(REFERENCE_ROOT, "models", "__init__.pyc")))


def sources_available():
    """True only in the build container (tests that read reference .py files / yaml configs)."""
    return os.path.exists(os.path.join(REFERENCE_ROOT, "models", "__init__.py"))


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
    repo = os.path.dirname(_HERE)
    if repo not in sys.path:  # workdir() changes the cwd while models are built: keep the repo importable
        sys.path.insert(1, repo)
    _installed = True


BERT_BASE_CONFIG = {
    "architectures": ["BertForMaskedLM"], "attention_probs_dropout_prob": 0.1, "hidden_act": "gelu",
    "hidden_dropout_prob": 0.1, "hidden_size": 768, "initializer_range": 0.02, "intermediate_size": 3072,
    "layer_norm_eps": 1e-12, "max_position_embeddings": 512, "model_type": "bert", "num_attention_heads": 12,
    "num_hidden_layers": 12, "pad_token_id": 0, "type_vocab_size": 2, "vocab_size": 30522,
}

# data/bert-large-uncased-12l of configs/pretrain/x2vlm_large_*.yaml: BERT-large width, 12 layers
BERT_LARGE_12L_CONFIG = dict(BERT_BASE_CONFIG, hidden_size=1024, intermediate_size=4096, num_attention_heads=16)


def workdir():
    """Scratch cwd holding configs/config_beit2_*.json (copied) and a synthesised data/bert-base-uncased."""
    global _workdir
    if _workdir is None:
        d = tempfile.mkdtemp(prefix="x2vlm_oracle_")
        os.makedirs(os.path.join(d, "configs"))
        # configs/config_beit2_{base,large}.json hold three values each (checkpoint path, width, patch size); written
        # here rather than copied so the GPU box (byte-code-only reference) has them too
        for n, width in (("base", 768), ("large", 1024)):
            with open(os.path.join(d, "configs", "config_beit2_%s.json" % n), "w") as fo:
                json.dump({"ckpt": "data/beitv2_%s_patch16_224_pt1k_ft21k.pth" % n, "vision_width": width, "patch_size": 16}, fo)
        vocab = ["[unused%d]" % i for i in range(30522)]
        vocab[0], vocab[100], vocab[101], vocab[102], vocab[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
        for name, cfg in (("bert-base-uncased", BERT_BASE_CONFIG), ("bert-large-uncased-12l", BERT_LARGE_12L_CONFIG)):
            bd = os.path.join(d, "data", name)
            os.makedirs(bd)
            with open(os.path.join(bd, "config.json"), "w") as f:
                json.dump(cfg, f)
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
    """1-rank process group: the reference's losses call dist.get_rank() / all_gather unconditionally
    (models/xvlm.py:805-806).  gloo serves CPU tensors, NCCL the CUDA ones when a GPU is present."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        import socket
        with socket.socket() as sk:  # a free port of our own: never the inherited MASTER_PORT (a parent may hold it)
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend, init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)


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


# ---------------------------------------------------------------------------------------------------------------
# the reference's UNMODIFIED callers on the B200-native modules (SURVEY.md §7.1 patch point (ii))
# ---------------------------------------------------------------------------------------------------------------
class x2k_patched:
    """Context manager: while active, the reference's model factories resolve to x2vlm_b200's drop-in modules —
    `models.beit2.{beit_base_patch16, beit_large_patch16, interpolate_pos_embed, load_pretrained_beit2}`,
    `models.xvlm.{BertModel, BertForMaskedLM}`, `models.model_generation.BertLMHeadModel`,
    `models.xbert.BertOnlyMLMHead` — so `XVLMBase.__init__` (models/xvlm.py:246-316) builds its encoders from them.
    No reference file is touched; on exit the reference's own classes are back, so one process can hold the same
    reference class twice: once on its own encoders, once on the x2k ones."""

    TARGETS = (("models.beit2", ("beit_base_patch16", "beit_large_patch16", "interpolate_pos_embed", "load_pretrained_beit2"), "beit2"),
               ("models.xvlm", ("BertModel", "BertForMaskedLM"), "xbert"),
               ("models.model_generation", ("BertLMHeadModel",), "xbert"),
               ("models.xbert", ("BertOnlyMLMHead",), "xbert"))

    def __enter__(self):
        import importlib
        install()
        from x2vlm_b200 import beit2, xbert
        mine = {"beit2": beit2, "xbert": xbert}
        self._saved = []
        for modname, names, src in self.TARGETS:
            mod = importlib.import_module(modname)
            for n in names:
                self._saved.append((mod, n, getattr(mod, n)))
                setattr(mod, n, getattr(mine[src], n))
        return self

    def __exit__(self, *exc):
        for mod, n, v in self._saved:
            setattr(mod, n, v)


def build_reference_model(cls_path="models.model_pretrain.XVLM", config=None, seed=0, x2k=False, **ctor_kw):
    """Construct a reference model class (by dotted path) — on the reference's own encoders, or with x2k=True on the
    B200-native drop-in modules (the class itself and everything it calls stay the reference's code)."""
    import contextlib
    import importlib
    import torch
    install()
    init_dist()
    cwd = os.getcwd()
    os.chdir(workdir())
    try:
        modname, clsname = cls_path.rsplit(".", 1)
        with (x2k_patched() if x2k else contextlib.nullcontext()):
            cls = getattr(importlib.import_module(modname), clsname)
            torch.manual_seed(seed)
            if not ctor_kw and clsname in ("XVLM",):
                ctor_kw = dict(load_vision_params=False, load_text_params=False, pretraining=False)
            m = cls(config=config or base_config(), **ctor_kw)
    finally:
        os.chdir(cwd)
    return m


def run_reference_retrieval(model, images, text_ids, text_atts, k_test, device, image_bs=64, text_bs=256):
    """The reference's own Retrieval.evaluation loop (Retrieval.py:71-157) on pre-tokenised tensors: the data loader and
    the tokenizer are stand-ins that hand out slices of `images` / `text_ids` / `text_atts`; everything that computes is
    the reference's code.  Returns (score_matrix_i2t, score_matrix_t2i) as numpy arrays, like the reference."""
    import importlib
    import torch
    from types import SimpleNamespace
    install()
    cwd = os.getcwd()
    os.chdir(workdir())
    try:
        Retrieval = importlib.import_module("Retrieval")
    finally:
        os.chdir(cwd)
    n_img, n_txt = images.shape[0], text_ids.shape[0]

    class Tok:
        def __call__(self, text, **kw):
            sel = torch.as_tensor(list(text), device=text_ids.device)
            out = SimpleNamespace(input_ids=text_ids[sel], attention_mask=text_atts[sel])
            out.to = lambda dev: SimpleNamespace(input_ids=out.input_ids.to(dev), attention_mask=out.attention_mask.to(dev))
            return out

    class Loader:
        dataset = SimpleNamespace(text=list(range(n_txt)), image=list(range(n_img)))

        def __iter__(self):
            for i in range(0, n_img, image_bs):
                yield images[i:i + image_bs], torch.arange(i, min(n_img, i + image_bs))

    Retrieval.args = SimpleNamespace(distributed=False)
    return Retrieval.evaluation(model, Loader(), Tok(), torch.device(device),
                                {"batch_size_test_text": text_bs, "max_tokens": text_ids.shape[1], "k_test": k_test})
