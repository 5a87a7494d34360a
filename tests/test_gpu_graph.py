"""Whole-step CUDA-graph replay (accelerator.GraphedStep): replays must draw FRESH dropout masks although every kernel
argument is frozen in the graph — the kernels add the device counter ops.dropout_base to their Philox offset and the
graph advances it.  Checked bit-exactly against the numpy Philox oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_step_advances_dropout_offsets():
    from oracle import philox
    from x2vlm_b200 import accelerator, ops
    from x2vlm_b200 import functional as XF
    dev = torch.device("cuda:0")
    M, N, K, p = 128, 256, 64, 0.5
    g = torch.Generator(device=dev).manual_seed(3)
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    Bm = torch.randn(N, K, device=dev, generator=g).bfloat16()
    ref = A.float() @ Bm.float().t()
    XF.manual_seed(5)
    base = ops.dropout_base(dev)
    base.zero_()
    seen = []

    def fn(inp):
        seed, off = XF.dropout_state.take(M * N)
        seen.append((seed, off))
        out = torch.empty(M, N, device=dev)
        ops.gemm(inp["a"], Bm, M, N, K, dropout_p=p, dropout_seed=seed, dropout_offset=off, out_f32=out)
        return out

    gs = accelerator.GraphedStep(fn, {"a": A}, warmup=2)
    per = gs.dropout_offsets_per_step
    assert per > 0 and gs.x2k_launches_per_step == 1
    seed, off_cap = seen[-1]  # the offset baked into the captured kernel node
    outs = [gs({"a": A}).clone() for _ in range(3)]
    torch.cuda.synchronize()
    assert int(base.item()) == 3 * per
    for r, o in enumerate(outs):
        keep = torch.from_numpy(philox.keep_scale(seed, off_cap + r * per, M * N, p)).view(M, N).to(dev)
        assert torch.equal(o == 0, keep == 0) or ((o == 0) != (keep == 0)).sum().item() <= 2  # exact zeros of ref are rare
        assert (o - ref * keep).abs().max().item() <= 1e-3 * ref.abs().max().item()
    assert not torch.equal(outs[0] == 0, outs[1] == 0)
    # eager launches after the replays keep drawing disjoint ranges (host counter + device base)
    eager = fn({"a": A})
    s2, off2 = seen[-1]
    keep = torch.from_numpy(philox.keep_scale(s2, off2 + 3 * per, M * N, p)).view(M, N).to(dev)
    assert (eager - ref * keep).abs().max().item() <= 1e-3 * ref.abs().max().item()
    base.zero_()


def test_graphed_step_staged_inputs():
    """stage() / run_staged(): the next batch's pinned host-to-device copy runs on a copy stream while the graph
    replays; results equal the direct call on the same inputs, batch after batch."""
    from x2vlm_b200 import accelerator, ops
    dev = torch.device("cuda:0")
    M, N, K = 256, 256, 128
    g = torch.Generator().manual_seed(1)
    Bm = torch.randn(N, K, generator=g).bfloat16().to(dev)

    def fn(inp):
        out = torch.empty(M, N, device=dev)
        ops.gemm(inp["a"], Bm, M, N, K, out_f32=out)
        return out

    host = [torch.randn(M, K, generator=g).bfloat16().pin_memory() for _ in range(4)]
    gs = accelerator.GraphedStep(fn, {"a": host[0].to(dev)}, warmup=1)
    gs.stage({"a": host[0]})
    for i in range(4):
        out = gs.run_staged()
        if i + 1 < 4:
            gs.stage({"a": host[i + 1]})       # overlaps the replay above
        got = out.clone()
        want = host[i].to(dev).float() @ Bm.float().t()
        assert (got - want).abs().max().item() <= 1e-3 * want.abs().max().item(), i
    gs.release()
