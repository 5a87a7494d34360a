"""Parameter plumbing for the x2k kernels: bf16 shadow weights, gradient sinks and the flat arena.

The drop-in modules keep ordinary fp32 ``nn.Parameter``s under the reference's names (the state_dict
contract).  The tensor-core GEMMs consume bf16 copies:

* ``Shadow`` — the bf16 copy of one weight, or of several weights packed row-wise (BERT's separate
  query/key/value Linears, models/xbert.py:234-241, run as ONE [3D, D] GEMM).  Stand-alone it
  re-casts a weight only when its ``_version`` changed.
* ``ParamArena`` — the B200-first layout for training: every parameter of a model is re-pointed into
  one flat fp32 buffer, with a flat fp32 gradient buffer (``param.grad`` are views) and a flat bf16
  shadow.  Packed groups are laid out adjacently, so a Shadow is a plain view; wgrad GEMMs accumulate
  straight into the flat gradient (no per-parameter grad tensors, no bucket copies), the optimizer is
  one fused kernel over the flat buffers (x2k_adamw_flat) and DDP all-reduces contiguous slices.
"""
import torch

from . import ops

_ALIGN = 8  # elements: 32 B in fp32, 16 B in bf16 (TMA needs 16-byte aligned global addresses)


class Shadow:
    """bf16 shadow + gradient sink of one or more [rows_i, K] fp32 parameters packed along dim 0."""

    def __init__(self, *params, K=None):
        self.params = list(params)
        self.K = K if K is not None else (params[0].shape[-1] if params[0].dim() > 1 else params[0].numel())
        self.rows = [p.numel() // self.K for p in params]
        self._buf = None
        self._versions = None
        self.arena = None  # set by ParamArena
        self.offset = None

    @property
    def total_rows(self):
        return sum(self.rows)

    def get_nograd(self):
        """Same as get() but never counted as a forward use (for backward passes)."""
        return self.get(count_use=False)

    def get(self, count_use=True):
        """bf16 [total_rows, K], current."""
        if self.arena is not None:
            if tuple(p._version for p in self.params) != self._versions:
                self.arena.sync_bf16()  # some parameter was modified in place since the last sync
            return self.arena.bf16[self.offset:self.offset + self.total_rows * self.K].view(self.total_rows, self.K)
        p0 = self.params[0]
        if self._buf is None or self._buf.device != p0.device:
            self._buf = torch.empty(self.total_rows, self.K, dtype=torch.bfloat16, device=p0.device)
            self._versions = [None] * len(self.params)
        r = 0
        for i, p in enumerate(self.params):
            ver = (p._version, p.data_ptr())
            if self._versions[i] != ver:
                ops.cast_f32_bf16(p.detach().contiguous().view(-1), self._buf[r:r + self.rows[i]].view(-1))
                self._versions[i] = ver
            r += self.rows[i]
        return self._buf

    def grad_sink(self):
        """fp32 [total_rows, K] view of the arena's flat gradient (accumulate into it), or None."""
        if self.arena is None:
            return None
        return self.arena.grad[self.offset:self.offset + self.total_rows * self.K].view(self.total_rows, self.K)

    def get_f32(self):
        """fp32 [total_rows * K] view of the packed parameters inside the arena (None without an arena)."""
        if self.arena is None:
            return None
        return self.arena.flat[self.offset:self.offset + self.total_rows * self.K]

    def grads_done(self):
        if self.arena is not None:
            self.arena.note_grad_written(self, self.params)

    def split_grad(self, g):
        """Split a packed [total_rows, K] gradient into per-parameter pieces (non-arena path)."""
        out, r = [], 0
        for p, n in zip(self.params, self.rows):
            out.append(g[r:r + n].view(p.shape))
            r += n
        return out


class GappedBias(Shadow):
    """BEiT's QKV bias (models/beit2.py:129): cat(q_bias, zeros, v_bias) — K has no bias parameter.  Inside an arena the
    two parameters are laid out with a zero gap of one bias length between them, so the [3D] bias the packed QKV GEMM adds
    is a plain view of the flat buffer (no torch.cat / zeros_like per block per forward) and the two bias gradients are
    column sums written straight into their slices of the flat gradient."""

    def __init__(self, q_bias, v_bias):
        super().__init__(q_bias, v_bias, K=q_bias.numel())
        self.gap = q_bias.numel()  # elements between the first and the second member

    def get_f32(self):
        if self.arena is None:
            return None
        return self.arena.flat[self.offset:self.offset + 3 * self.K]

    def grad_views(self):
        g = self.arena.grad
        return g[self.offset:self.offset + self.K], g[self.offset + 2 * self.K:self.offset + 3 * self.K]

    def get(self, count_use=True):
        raise NotImplementedError("GappedBias has no bf16 shadow: the GEMM epilogue reads the fp32 bias")


def collect_shadows(model):
    seen, out = set(), []
    for m in model.modules():
        refresh = getattr(m, "_x2k_refresh", None)
        if refresh is not None:
            refresh()  # modules whose weight is re-pointed after construction (tied MLM decoder)
        for s in getattr(m, "_x2k_shadows", ()):
            if id(s) not in seen:
                seen.add(id(s))
                out.append(s)
    return out


class ParamArena:
    """Flat fp32 params / fp32 grads / bf16 shadow for every parameter of `model` (on its device)."""

    def __init__(self, model):
        params = [p for p in model.parameters()]
        if not params:
            raise ValueError("model has no parameters")
        dev = params[0].device  # (a CPU arena is host logic only: sync_bf16 needs the CUDA cast kernel)
        self.model = model
        shadows = collect_shadows(model)
        order, placed = [], set()
        # module registration order, except that the members of a packed group follow its first member
        group_of = {}
        for s in shadows:
            for p in s.params:
                group_of.setdefault(id(p), s)
        for p in params:
            if id(p) in placed:
                continue
            s = group_of.get(id(p))
            members = s.params if s is not None and len(s.params) > 1 else [p]
            for q in members:
                if id(q) not in placed:
                    placed.add(id(q))
                    order.append(q)
        self.params = order
        self.offsets = {}
        off = 0
        for p in order:
            s = group_of.get(id(p))
            packed_tail = s is not None and len(s.params) > 1 and p is not s.params[0]
            if not packed_tail:
                off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
            elif p is s.params[1]:
                off += getattr(s, "gap", 0)  # zero-filled hole after the group's first member (GappedBias)
            self.offsets[id(p)] = off
            off += p.numel()
        self.numel = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.bf16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=dev)
        with torch.no_grad():
            for p in order:
                o = self.offsets[id(p)]
                self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
                p._x2k_grad, p._x2k_arena = p.grad, self  # fused backwards accumulate small gradients in place
        for s in shadows:
            o = self.offsets[id(s.params[0])]
            for a, b in zip(s.params[:-1], s.params[1:]):
                gap = getattr(s, "gap", 0) if a is s.params[0] else 0
                assert self.offsets[id(b)] == self.offsets[id(a)] + a.numel() + gap, "packed group is not adjacent"
            s.arena, s.offset = self, o
        self.shadows = shadows
        self._pending = {}      # id(shadow) -> outstanding uses in the current backward
        self.on_grad_ready = None  # callback(list of params) — installed by the DDP accelerator
        self._hooks = []
        sink_params = {id(p) for s in shadows for p in s.params}
        # parameters whose gradient (also) arrives through ordinary autograd accumulation
        self.autograd_params = [p for p in order if p.requires_grad and
                                (id(p) not in sink_params or getattr(p, "_x2k_autograd_too", False))]
        self.sink_param_ids = sink_params
        for p in self.autograd_params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._autograd_hook))
        self.mark_dirty()

    # -- bf16 shadow upkeep -------------------------------------------------------------------
    def mark_dirty(self):
        for s in self.shadows:
            s._versions = None

    def sync_bf16(self):
        """Re-cast the whole flat fp32 buffer into the bf16 shadow and snapshot parameter versions.
        (x2k_adamw_flat writes both buffers itself and does not bump versions, so a training loop
        never comes back here after the first step.)"""
        ops.cast_f32_bf16(self.flat, self.bf16)
        for s in self.shadows:
            s._versions = tuple(p._version for p in s.params)

    # -- gradient bookkeeping -----------------------------------------------------------------
    def zero_grad(self):
        # FlatAdamW.step leaves the gradient buffer zeroed (x2k_adamw_flat, zero_grad = 1): the 1 GB fill is skipped then
        if not getattr(self, "grad_is_zero", False):
            self.grad.zero_()
        self.grad_is_zero = False
        self._pending.clear()

    def note_use(self, obj):
        """One forward use of a shadow / parameter whose gradient a fused backward will write in place."""
        self._pending[id(obj)] = self._pending.get(id(obj), 0) + 1

    def note_grad_written(self, obj, params):
        """One backward contribution has been accumulated; after the last one of this step the parameters are final."""
        n = self._pending.get(id(obj), 1) - 1
        self._pending[id(obj)] = n
        if n <= 0 and self.on_grad_ready is not None:
            self.on_grad_ready(params)

    def _autograd_hook(self, p):
        if self.on_grad_ready is not None:
            self.on_grad_ready([p])

    def span(self, p):
        o = self.offsets[id(p)]
        return o, o + p.numel()
