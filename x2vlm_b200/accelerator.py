"""Data-parallel runtime for the flat-arena model: bucketed NCCL all-reduce overlapped with backward,
fused global-norm clipping and a one-kernel AdamW.

`X2kDDPAccelerator` keeps the four-method surface of the reference's
accelerators/apex_ddp_accelerator.py:ApexDDPAccelerator (`set_up`, `broadcast`, `backward_step`,
`optimizer_step`; called from Pretrain.py:576-578, :67-72) and the wrapped model exposes `.module`
(Pretrain.py:328).  Differences, all deliberate:

* apex DDP with ``delay_allreduce=True`` issues ONE un-overlapped all-reduce of 1.02 GB after the last
  gradient.  Here the flat fp32 gradient buffer of `ParamArena` is cut into contiguous
  buckets; a bucket is all-reduced (SUM then 1/W, apex's ``gradient_average``) on a side stream the
  moment its last gradient has been written, so communication hides under the rest of backward.
  The wgrad GEMMs accumulate directly into that buffer — there are no per-parameter gradient
  tensors and no bucket copies.
* bf16 needs no loss scaler, so `backward_step` is a plain backward.
* `optimizer_step` computes the global gradient norm on the device and leaves the clip coefficient
  in device memory for the fused AdamW kernel (no host sync).
"""
import os

import torch
import torch.distributed as dist

from . import ops
from .params import ParamArena

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight",
            "norm2.bias", "norm2.weight")


def name_based_groups(model, params, lr=1e-4, weight_decay=0.01, lr_mult=1.0, vision_lr=None, text_lr=None, cross_lr=None):
    """The reference's parameter groups (optim.py:26-104): decay / no-decay x {default, init_params * lr_mult, vision, text,
    cross}, in the reference's group order; frozen parameters are left out."""
    names = {id(p): n for n, p in model.named_parameters()}
    init_params = set(getattr(model, "init_params", []) or [])
    if cross_lr is None:
        cross_lr = text_lr
    groups = [{"params": [], "weight_decay": weight_decay, "lr": lr}, {"params": [], "weight_decay": 0.0, "lr": lr},
              {"params": [], "weight_decay": weight_decay, "lr": lr * lr_mult}, {"params": [], "weight_decay": 0.0, "lr": lr * lr_mult}]
    if vision_lr is not None:
        for lr_ in (vision_lr, text_lr, cross_lr):
            groups += [{"params": [], "weight_decay": weight_decay, "lr": lr_}, {"params": [], "weight_decay": 0.0, "lr": lr_}]
    for p in params:
        n = names.get(id(p), "")
        if not p.requires_grad:
            continue
        nd = int(any(k in n for k in NO_DECAY))
        if vision_lr is not None and n.startswith("vision_encoder"):
            gi = 4
        elif vision_lr is not None and n.startswith("text_encoder"):
            gi = 6
        elif vision_lr is not None and n.startswith("cross_encoder"):
            gi = 8
        elif n in init_params:
            gi = 2
        else:
            gi = 0
        groups[gi + nd]["params"].append(p)
    return groups


class FlatAdamW(torch.optim.Optimizer):
    """AdamW over the arena's flat buffers — one fused kernel (x2k_adamw_flat) for every parameter.

    A real `torch.optim.Optimizer`: `param_groups` hold per-group lr / weight_decay / betas / eps (so LambdaLR and the
    reference's `reinit_scheduler_properties_mysched`, Pretrain.py:33-51, drive it like any optimizer), `state_dict` /
    `load_state_dict` carry the flat moments and the step count (checkpoint `training_states`, Pretrain.py:603).
    Built either from the reference's name-based rules (optim.py:26-104) or from the param_groups of an existing torch
    AdamW (`from_optimizer`): the groups, their order and their hyper-parameters are taken over one to one."""

    def __init__(self, model, arena, lr=1e-4, weight_decay=0.01, lr_mult=1.0, vision_lr=None, text_lr=None, cross_lr=None,
                 betas=(0.9, 0.98), eps=1e-8, param_groups=None):
        self.arena = arena
        if param_groups is None:
            param_groups = name_based_groups(model, arena.params, lr, weight_decay, lr_mult, vision_lr, text_lr, cross_lr)
        known = {id(p) for p in arena.params}
        for g in param_groups:
            for p in g["params"]:
                if id(p) not in known:
                    raise ValueError("FlatAdamW: optimizer parameter is not part of the model's arena")
        if not any(g["params"] for g in param_groups):
            raise ValueError("FlatAdamW: no trainable parameters")
        super().__init__(param_groups, dict(lr=lr, weight_decay=weight_decay, betas=tuple(betas), eps=eps))
        b, e = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        if any(tuple(g["betas"]) != tuple(b) or g["eps"] != e for g in self.param_groups):
            raise NotImplementedError("FlatAdamW: betas / eps must be the same in every group (they are kernel constants)")
        self.betas, self.eps = tuple(b), e
        for g in self.param_groups:
            g.setdefault("initial_lr", g["lr"])
        dev = arena.flat.device
        # one segment per parameter (arena order == ascending offsets); parameters outside every group are frozen
        seg_end, seg_group = [], []
        gid = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        for p in arena.params:
            seg_end.append(arena.span(p)[1])
            seg_group.append(gid.get(id(p), -1))
        seg_end[-1] = arena.numel
        n_groups = len(self.param_groups)
        self.seg_end = torch.tensor(seg_end, dtype=torch.int64, device=dev)
        self.seg_lr = torch.zeros(len(seg_end), dtype=torch.float32, device=dev)
        self.seg_wd = torch.zeros(len(seg_end), dtype=torch.float32, device=dev)
        self._seg_group_t = torch.tensor([g if g >= 0 else n_groups for g in seg_group], device=dev)
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.grad_scale = torch.ones(1, dtype=torch.float32, device=dev)
        self._lr_host = None
        self._upload_hparams()

    @classmethod
    def from_optimizer(cls, model, arena, optimizer):
        """Take over a torch AdamW (e.g. the reference's optim.create_optimizer result): same groups, lr, weight decay,
        betas, eps.  Anything that is not a decoupled-weight-decay Adam is rejected loudly."""
        if isinstance(optimizer, FlatAdamW):
            return optimizer
        if not isinstance(optimizer, torch.optim.AdamW):
            raise TypeError("X2kDDPAccelerator needs a torch.optim.AdamW (the reference's optim.create_optimizer) or None, got %s"
                            % type(optimizer).__name__)
        if any(g.get("amsgrad", False) or g.get("maximize", False) for g in optimizer.param_groups):
            raise NotImplementedError("FlatAdamW: amsgrad / maximize are not implemented")
        if any(len(st) for st in optimizer.state.values()):
            raise NotImplementedError("FlatAdamW.from_optimizer: the incoming optimizer already has state; load it with "
                                      "FlatAdamW.load_state_dict instead")
        groups = [{"params": list(g["params"]), "lr": g["lr"], "weight_decay": g["weight_decay"], "betas": tuple(g["betas"]),
                   "eps": g["eps"], **({"initial_lr": g["initial_lr"]} if "initial_lr" in g else {})}
                  for g in optimizer.param_groups]
        d = optimizer.defaults
        return cls(model, arena, lr=d["lr"], weight_decay=d["weight_decay"], betas=d["betas"], eps=d["eps"], param_groups=groups)

    def _upload_hparams(self):
        lrs = [float(g["lr"]) for g in self.param_groups] + [0.0]   # frozen parameters: lr 0, wd 0
        wds = [float(g["weight_decay"]) for g in self.param_groups] + [0.0]
        if self._lr_host == (lrs, wds):
            return
        self._lr_host = (list(lrs), list(wds))
        dev = self.seg_lr.device
        self.seg_lr.copy_(torch.tensor(lrs, dtype=torch.float32, device=dev)[self._seg_group_t])
        self.seg_wd.copy_(torch.tensor(wds, dtype=torch.float32, device=dev)[self._seg_group_t])

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()  # the gradients are views of one flat buffer: one fill, never set to None

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FlatAdamW.step: closures are not supported")
        self._upload_hparams()  # picks up LambdaLR-style mutations of param_groups[i]['lr']
        self.step_dev += 1
        a = self.arena
        # (x2k_adamw_flat can also zero the gradients it consumed — zero_grad=1 — but the ninth memory stream drops the
        #  kernel from 5.7 to 1.5 TB/s on B200: 5.8 ms instead of 1.3 ms + a 0.3 ms fill, profiles/r02g_launches_step.md)
        ops.adamw_flat(a.flat, a.grad, self.exp_avg, self.exp_avg_sq, a.bf16, a.numel, self.seg_end, self.seg_lr, self.seg_wd,
                       self.betas[0], self.betas[1], self.eps, step_dev=self.step_dev, grad_scale=self.grad_scale)
        self.grad_scale.fill_(1.0)

    def state_dict(self):
        return {"flat": True, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": int(self.step_dev.item()),
                "numel": self.arena.numel,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        """Resume from `state_dict()` (the checkpoint's `training_states`): flat moments, step count, group hyper-parameters."""
        if not sd.get("flat") or sd["numel"] != self.arena.numel or len(sd["param_groups"]) != len(self.param_groups):
            raise ValueError("FlatAdamW.load_state_dict: not a FlatAdamW state of this model (numel %s vs %d, %d vs %d groups)"
                             % (sd.get("numel"), self.arena.numel, len(sd.get("param_groups", ())), len(self.param_groups)))
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_dev.fill_(int(sd["step"]))
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "params"})
        self._lr_host = None
        self._upload_hparams()


def rebind_scheduler(lr_scheduler, optimizer):
    """The scheduler the caller built on its torch optimizer, re-created on `optimizer` (same lambdas, same progress).
    Only LambdaLR — the one scheduler the reference creates (scheduler.py:4-33) — is supported."""
    if lr_scheduler is None or getattr(lr_scheduler, "optimizer", None) is optimizer:
        return lr_scheduler
    from torch.optim.lr_scheduler import LambdaLR
    if not isinstance(lr_scheduler, LambdaLR):
        raise TypeError("X2kDDPAccelerator.set_up: only LambdaLR schedulers can be re-bound to the flat optimizer, got %s"
                        % type(lr_scheduler).__name__)
    lambdas = list(lr_scheduler.lr_lambdas)
    if len(lambdas) != len(optimizer.param_groups):
        raise ValueError("scheduler has %d lr_lambdas, optimizer %d groups" % (len(lambdas), len(optimizer.param_groups)))
    done = lr_scheduler.last_epoch
    for g in optimizer.param_groups:  # LambdaLR multiplies the groups' initial_lr
        g["lr"] = g.get("initial_lr", g["lr"])
    new = LambdaLR(optimizer, lambdas, last_epoch=-1)
    for _ in range(max(0, done)):  # replay the progress already made (0 for a fresh run)
        new.step()
    return new


class DDPModel(torch.nn.Module):
    """Thin wrapper exposing `.module` like apex / torch DDP."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


class GradBucketer:
    """Contiguous buckets over the arena's flat gradient; all-reduce each as soon as it is complete."""

    def __init__(self, arena, world_size, bucket_mb=48.0, process_group=None):
        self.arena, self.world_size, self.pg = arena, world_size, process_group
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets = []  # [start, end, [params]]
        cur = None
        for p in arena.params:
            if not p.requires_grad:
                continue
            s, e = arena.span(p)
            if cur is None or (e - cur[0]) > cap:
                cur = [s, e, []]
                self.buckets.append(cur)
            cur[1] = e
            cur[2].append(p)
        self.bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b[2]}
        # notifications needed before a parameter's gradient is final in one backward
        self.need = {}
        auto = {id(p) for p in arena.autograd_params}
        for p in arena.params:
            if p.requires_grad:
                self.need[id(p)] = int(id(p) in arena.sink_param_ids) + int(id(p) in auto)
        self.is_cuda = arena.grad.is_cuda
        self.comm_stream = torch.cuda.Stream() if self.is_cuda else None
        self.launch_order = []
        self.reset()
        arena.on_grad_ready = self.on_grad_ready

    def reset(self):
        self.remaining = dict(self.need)
        self.bucket_left = [len(b[2]) for b in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.launch_order = []

    def on_grad_ready(self, params):
        for p in params:
            k = id(p)
            if k not in self.remaining:
                continue
            self.remaining[k] -= 1
            if self.remaining[k] == 0:
                bi = self.bucket_of[k]
                self.bucket_left[bi] -= 1
                if self.bucket_left[bi] == 0:
                    self._launch(bi)

    def _launch(self, bi):
        if self.launched[bi]:
            return
        self.launched[bi] = True
        self.launch_order.append(bi)
        if self.world_size <= 1:
            return
        s, e, _ = self.buckets[bi]
        view = self.arena.grad[s:e]
        if self.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                # NCCL averages inside the collective (apex gradient_average): no second pass over the bucket
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.pg)
        else:  # gloo (CPU tests of the bucket logic)
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg)
            view.mul_(1.0 / self.world_size)

    def finalize(self):
        """Reduce whatever did not complete (parameters without gradient this step) and join the streams."""
        for bi in range(len(self.buckets)):
            self._launch(bi)
        if self.is_cuda and self.world_size > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        order = self.launch_order
        self.reset()
        return order


def _flatten(tree, prefix=""):
    """{path: tensor} over nested dicts / lists / tuples of tensors (None leaves are skipped)."""
    out = {}
    if isinstance(tree, torch.Tensor):
        out[prefix] = tree
    elif isinstance(tree, dict):
        for k, v in tree.items():
            out.update(_flatten(v, prefix + "/" + str(k)))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out.update(_flatten(v, prefix + "/" + str(i)))
    elif tree is not None:
        raise TypeError("GraphedStep inputs must be tensors or dicts/lists of tensors, got %r" % type(tree))
    return out


class GraphedStep:
    """A whole training step — zero_grad, forward, backward (with the bucketed all-reduce), clip, AdamW — captured
    ONCE into a CUDA graph and replayed: ~8 000 kernel launches per step cost the host one cudaGraphLaunch instead
    of ~50 ms of Python/ctypes enqueue work (the step is launch-bound otherwise: bench.py reports
    `host_enqueue_ms_per_step`).

    `fn(inputs)` must be sync-free (no .item(), no pageable H2D copies) and return a tensor or a dict of tensors.
    Inputs are copied into static device buffers before each replay, outputs are returned as static tensors (valid
    until the next call).  What changes between replays although kernel arguments are frozen:
      * the in-kernel dropout masks — every kernel adds the device counter `ops.dropout_base` to its Philox offset and
        the graph's last node advances it by the offsets one step consumes;
      * torch's own RNG consumers (hard-negative multinomial, DropPath) — torch.cuda.graph registers the default
        generator, which is advanced per replay;
      * the AdamW step count (a device tensor) and the learning rates (seg_lr device tensor, refreshed from
        `optimizer.param_groups` before each replay).
    """

    def __init__(self, fn, example_inputs, optimizer=None, warmup=3):
        from . import functional as XF
        flat = _flatten(example_inputs)
        if not flat:
            raise ValueError("GraphedStep needs at least one input tensor")
        dev = next(iter(flat.values())).device
        self.device = dev
        self.fn, self.optimizer = fn, optimizer
        self.static_in = self._clone_tree(example_inputs)
        self._flat_in = _flatten(self.static_in)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # lazy inits (cudaFuncSetAttribute, allocator pools, autotuners) happen here
                fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        base = ops.dropout_base(dev)
        off0 = XF.dropout_state.offset
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)
            self.x2k_launches_per_step = ops.launch_count() - l0  # libx2k kernel nodes in the graph
            self.dropout_offsets_per_step = XF.dropout_state.offset - off0
            base.add_(self.dropout_offsets_per_step)
        self.replays = 0

    # -- input prefetch: the next batch's host-to-device copy overlaps the current replay ---------------------------
    def stage(self, inputs):
        """Start copying `inputs` (pinned host tensors) into a second set of device buffers on a copy stream; returns
        immediately.  `run_staged()` then moves them into the graph's static buffers (device-to-device) and replays."""
        if getattr(self, "_staging", None) is None:
            self._staging = self._clone_tree(self.static_in)
            self._flat_staging = _flatten(self._staging)
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staged_evt = torch.cuda.Event()
            self._consumed_evt = None
        if self._consumed_evt is not None:  # the previous batch has left the staging buffers
            self._copy_stream.wait_event(self._consumed_evt)
        with torch.cuda.stream(self._copy_stream):
            for path, t in _flatten(inputs).items():
                self._flat_staging[path].copy_(t, non_blocking=True)
            self._staged_evt.record(self._copy_stream)

    def run_staged(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged_evt)
        for path, t in self._flat_staging.items():
            self._flat_in[path].copy_(t, non_blocking=True)
        self._consumed_evt = torch.cuda.Event()
        self._consumed_evt.record(cur)
        return self.__call__(None)

    def release(self):
        """Destroy the captured graph (NCCL will not tear a communicator down while a graph still holds its kernels)."""
        self.static_out = None
        if self.graph is not None:
            self.graph.reset()
            self.graph = None

    @staticmethod
    def _clone_tree(tree):
        if isinstance(tree, torch.Tensor):
            return tree.clone()
        if isinstance(tree, dict):
            return {k: GraphedStep._clone_tree(v) for k, v in tree.items()}
        if isinstance(tree, (list, tuple)):
            return type(tree)(GraphedStep._clone_tree(v) for v in tree)
        return tree

    def __call__(self, inputs=None):
        if inputs is not None:
            for path, t in _flatten(inputs).items():
                dst = self._flat_in[path]
                if dst.data_ptr() != t.data_ptr():
                    dst.copy_(t, non_blocking=True)  # H2D from pinned memory or D2D; stream-ordered before the replay
        if self.optimizer is not None and hasattr(self.optimizer, "_upload_hparams"):
            self.optimizer._upload_hparams()
        self.graph.replay()
        self.replays += 1
        return self.static_out


class X2kDDPAccelerator:
    def __init__(self, cfg=None, logger=None):
        self.cfg = cfg or {}
        self.logger = logger
        self.arena = None
        self.bucketer = None
        self.norm_sq = None

    def set_up(self, model, optimizer, lr_scheduler, local_rank, world_size, rank):
        """Same call as ApexDDPAccelerator.set_up (accelerators/apex_ddp_accelerator.py:42-72; Pretrain.py:578): move the
        model to cuda:local_rank, join the NCCL group, flatten parameters, broadcast rank 0's weights, arm the bucketed
        all-reduce — and return (wrapped model, optimizer, lr_scheduler).

        The returned optimizer is always the flat fused AdamW: built from the incoming torch AdamW's param_groups (the
        reference's optim.create_optimizer result: groups, lr, weight decay, betas, eps are taken over one to one), or
        from self.cfg when `optimizer` is None; any other optimizer type is rejected.  The returned scheduler is the
        incoming LambdaLR re-created on that optimizer, so `reinit_scheduler_properties_mysched(optimizer, scheduler, ...)`
        and `scheduler.step()` keep working on what set_up returns."""
        torch.cuda.set_device(local_rank)
        model = model.cuda()
        # The gradient all-reduce moves ~1 GB per step under a ~25 ms backward: a few NCCL CTAs saturate that need, and
        # every SM NCCL takes is one the persistent GEMMs (one CTA pair per TPC, tiles dealt statically) then wait for.
        os.environ.setdefault("NCCL_MAX_CTAS", str(self.cfg.get("nccl_max_ctas", 8)))
        if world_size > 1 and not dist.is_initialized():
            addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
            port = int(os.environ.get("MASTER_PORT", 34171))
            dist.init_process_group(backend="nccl", init_method="tcp://{}:{}".format(addr, port), world_size=world_size,
                                    rank=rank)
        self.world_size = world_size
        self.arena = ParamArena(model)
        if world_size > 1:
            self.broadcast(model)
        self.bucketer = GradBucketer(self.arena, world_size, float(self.cfg.get("bucket_mb", 48.0)))
        if optimizer is None:
            optimizer = FlatAdamW(model, self.arena, **{k: v for k, v in self.cfg.items()
                                                        if k in ("lr", "weight_decay", "lr_mult", "vision_lr", "text_lr",
                                                                 "cross_lr")})
        else:
            optimizer = FlatAdamW.from_optimizer(model, self.arena, optimizer)
        lr_scheduler = rebind_scheduler(lr_scheduler, optimizer)
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=self.arena.flat.device)
        # spans of frozen parameters: kept at zero gradient so they never enter the clip norm (torch.clip_grad_norm_ skips them)
        self._frozen_spans = [self.arena.span(p) for p in self.arena.params if not p.requires_grad]
        return DDPModel(model), optimizer, lr_scheduler

    def broadcast(self, model, src=0):
        """One broadcast of the flat parameter buffer (the reference loops over 587 state_dict tensors)."""
        dist.broadcast(self.arena.flat, src)
        for b in model.buffers():
            dist.broadcast(b, src)
        self.arena.mark_dirty()

    def graph_step(self, fn, example_inputs, optimizer=None, warmup=3):
        """Capture `fn(inputs)` (a full sync-free training step) into a CUDA graph; see GraphedStep."""
        return GraphedStep(fn, example_inputs, optimizer=optimizer, warmup=warmup)

    def backward_step(self, loss, optimizer=None):
        loss.backward()
        return self.bucketer.finalize()

    def optimizer_step(self, optimizer, model, grad_norm):
        """Global-norm clip (torch.nn.utils.clip_grad_norm_ semantics) folded into the optimizer's gradient scale.
        Returns the total norm as a 0-dim device tensor."""
        if not isinstance(optimizer, FlatAdamW):
            raise TypeError("optimizer_step needs the optimizer that set_up returned")
        for s0, e0 in getattr(self, "_frozen_spans", ()):  # a packed group with one frozen member still receives a wgrad
            self.arena.grad[s0:e0].zero_()
        self.norm_sq.zero_()
        ops.sumsq(self.arena.grad, self.norm_sq, self.arena.numel)
        total_norm = self.norm_sq.sqrt()
        optimizer.grad_scale.copy_(torch.clamp(float(grad_norm) / (total_norm + 1e-6), max=1.0))
        return total_norm[0]
