"""GPU check of x2k_gemm against torch (run under gpurun). Prints one line per case."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x2vlm_b200 import _capi as C

L = C.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run_gemm(A, B, M, N, K, a_mn, b_mn, tile_n=0, **ep):
    args = C.X2kGemmArgs()
    args.A = A.data_ptr(); args.B = B.data_ptr()
    args.M, args.N, args.K = M, N, K
    args.lda = A.stride(0); args.ldb = B.stride(0)
    args.a_mn_major = a_mn; args.b_mn_major = b_mn
    args.tile_n = tile_n
    keep = []
    for k, v in ep.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
            setattr(args, k, v.data_ptr())
        else:
            setattr(args, k, v)
    rc = L.x2k_gemm(ctypes.byref(args), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    C.check(rc, "x2k_gemm")


def case(name, M, N, K, a_mn, b_mn, tile_n):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().t()
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    try:
        run_gemm(As, Bs, M, N, K, a_mn, b_mn, tile_n, out_f32=out, ld_out_f32=N)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print("CASE %-40s EXC %s" % (name, e)); return False
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= 2e-3 * scale + 1e-3
    print("CASE %-40s M=%d N=%d K=%d a_mn=%d b_mn=%d tn=%d  maxerr=%.4g (ref max %.3g) nan=%d %s" % (
        name, M, N, K, a_mn, b_mn, tile_n, err, scale, int(torch.isnan(out).sum()), "OK" if ok else "FAIL"))
    if not ok:
        d = (out - ref).abs()
        d = torch.nan_to_num(d, nan=1e9)
        blk = d[:min(M,128), :min(N,128)].reshape(-1, 8, min(N,128) // 16, 16).amax(dim=(1, 3))
        print("   err by 8x16 blocks (first 128x128):")
        for r in blk[:16]:
            print("   " + " ".join("%7.2g" % x for x in r.tolist()))
    return ok


def main():
    print("x2k version", L.x2k_version(), torch.cuda.get_device_name(0))
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    allok = True
    if mode in ("all", "basic"):
        for tn in (128, 256):
            allok &= case("kk-small", 128, 256, 64, 0, 0, tn)
            allok &= case("kk-k256", 256, 512, 256, 0, 0, tn)
            allok &= case("kk-tails", 200, 328, 200, 0, 0, tn)
    if mode in ("all", "mn"):
        for tn in (128, 256):
            allok &= case("k-mn (dgrad)", 256, 512, 256, 0, 1, tn)
            allok &= case("mn-k", 256, 512, 256, 1, 0, tn)
            allok &= case("mn-mn (wgrad)", 256, 512, 256, 1, 1, tn)
            allok &= case("mn-mn big", 768, 3072, 12608, 1, 1, tn)
    if mode in ("all", "epi"):
        allok &= epilogue_cases()
    if mode in ("all", "perf"):
        perf()
    print("ALL OK" if allok else "SOME FAILED")


def epilogue_cases():
    ok_all = True
    M, N, K = 394, 768, 768
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev); gamma = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev); rs = torch.rand(2, device=dev) + 0.5
    acc = A.float() @ W.float().t()
    # 1) bias + gelu + preact
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(o)
    run_gemm(A, W, M, N, K, 0, 0, 0, bias=bias, act=C.ACT_GELU, preact_out=pre, ld_preact=N, out_bf16=o, ld_out_bf16=N)
    ref_pre = acc + bias; ref = torch.nn.functional.gelu(ref_pre)
    e1 = (o.float() - ref).abs().max().item(); e2 = (pre.float() - ref_pre).abs().max().item()
    print("EPI bias+gelu+preact err %.4g %.4g" % (e1, e2)); ok_all &= e1 < 0.05 and e2 < 0.05
    # 2) bias, gamma, row_scale, residual -> fp32
    o32 = torch.empty(M, N, device=dev)
    run_gemm(A, W, M, N, K, 0, 0, 0, bias=bias, gamma=gamma, row_scale=rs, rows_per_scale=197, residual=res, ld_res=N,
             out_f32=o32, ld_out_f32=N, preact_out=pre, ld_preact=N)
    rsx = rs.repeat_interleave(197)[:, None]
    ref = res + (acc + bias) * gamma * rsx
    e = (o32 - ref).abs().max().item(); print("EPI layerscale+droppath+residual err %.4g" % e); ok_all &= e < 2e-3 * ref.abs().max().item()
    # 3) gelu bwd
    h = torch.randn(M, N, device=dev).bfloat16()
    run_gemm(A, W, M, N, K, 0, 0, 0, act=C.ACT_GELU_BWD, aux=h, ld_aux=N, out_bf16=o, ld_out_bf16=N)
    hf = h.float().requires_grad_(True); torch.nn.functional.gelu(hf).sum().backward()
    ref = acc * hf.grad
    e = (o.float() - ref).abs().max().item(); print("EPI gelu_bwd err %.4g (max %.3g)" % (e, ref.abs().max().item())); ok_all &= e < 0.02 * ref.abs().max().item()
    # 4) accumulate
    o32b = o32.clone()
    run_gemm(A, W, M, N, K, 0, 0, 0, accumulate=1, out_f32=o32b, ld_out_f32=N)
    e = (o32b - (o32 + acc)).abs().max().item(); print("EPI accumulate err %.4g" % e); ok_all &= e < 1e-2
    # 5) dropout: keep fraction and scaling
    run_gemm(A, W, M, N, K, 0, 0, 0, dropout_p=0.1, dropout_seed=1234, dropout_offset=77, out_f32=o32, ld_out_f32=N)
    kept = (o32 != 0)
    frac = kept.float().mean().item()
    e = ((o32 - acc / 0.9) * kept).abs().max().item()
    print("EPI dropout keep frac %.4f (expect 0.9) err on kept %.4g" % (frac, e)); ok_all &= abs(frac - 0.9) < 0.01 and e < 1e-2
    return ok_all


def perf():
    shapes = [("qkv", 12608, 2304, 768), ("proj", 12608, 768, 768), ("fc1", 12608, 3072, 768), ("fc2", 12608, 768, 3072),
              ("text-qkv", 5120, 2304, 768), ("fus-ffn1", 10240, 3072, 768), ("8k", 8192, 8192, 8192)]
    for name, M, N, K in shapes:
        A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for tn in (128, 256):
            for _ in range(3):
                run_gemm(A, W, M, N, K, 0, 0, tn, out_bf16=o, ld_out_bf16=N)
            torch.cuda.synchronize()
            st = torch.cuda.Event(enable_timing=True); en = torch.cuda.Event(enable_timing=True)
            st.record()
            for _ in range(20):
                run_gemm(A, W, M, N, K, 0, 0, tn, out_bf16=o, ld_out_bf16=N)
            en.record(); torch.cuda.synchronize()
            ms = st.elapsed_time(en) / 20
            print("PERF %-9s %5dx%5dx%5d tn=%d  %.3f ms  %.1f TFLOP/s" % (name, M, N, K, tn, ms, 2.0 * M * N * K / ms / 1e9))
        for _ in range(3):
            torch.matmul(A, W.t(), out=o)
        torch.cuda.synchronize()
        st = torch.cuda.Event(enable_timing=True); en = torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(20):
            torch.matmul(A, W.t(), out=o)
        en.record(); torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 20
        print("PERF %-9s cublas           %.3f ms  %.1f TFLOP/s" % (name, ms, 2.0 * M * N * K / ms / 1e9))


if __name__ == "__main__":
    main()
