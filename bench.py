#!/usr/bin/env python
"""bench.py — X2VLM-base pre-training step (ITC + ITM + MLM + bbox) on synthetic data, bf16, N GPUs of one node.

    python bench.py --gpus 1 --steps K --warmup W            # this framework
    torchrun ... bench.py --gpus N --steps K --warmup W      # one rank per GPU, NCCL
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference on the host cores (oracle/_ref)

One "step" = Pretrain.py:run_mixed_iter semantics on one synthetic batch per GPU:
zero_grad -> image iteration (64 image-text pairs: ITC+ITM+MLM) + region iteration (64 region-text samples over
26 images: ITC+ITM+MLM+bbox) summed -> one backward (bucketed all-reduce overlapped) -> clip 1.0 -> AdamW.
pairs/step/GPU = 128.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs (fwd+bwd, 2 FLOP / MAC) of the reference algorithm, SURVEY.md §8d
# average DRAM bytes per GEMM launch of the step (dram__bytes_read.sum + dram__bytes_write.sum over the 423 tcgen05 GEMM
# launches of profiles/r02_final_launches_step.md; round 1: 124.37 MB over 422), from the committed ncu capture — not
# re-measured at bench time
ROOFLINE_TRAFFIC = 121.32e6
FLOP_IMAGE_PAIR = 231.4e9            # image iteration, per pair
FLOP_REGION_SAMPLE = 3 * 48.92e9     # region iteration, per region-text sample (text x2, fusion x5, head)
FLOP_VISION_IMAGE = 3 * 35.13e9      # + vision once per unique region image


# --config: the BASELINE.json workloads.  FLOPs per unit are the reference-algorithm closed forms of SURVEY.md §8(d)
# (fwd + bwd = 3 x fwd): image-iteration pair, region sample (2 text + 5 fusion + head), vision once per unique region image.
CONFIGS = {
    "base": dict(model="base", name="X2VLM-base", batch=64, region_images=26, flop_pair=231.4e9, flop_region=3 * 48.92e9,
                 flop_image=3 * 35.13e9, what="pretrain step (ITC+ITM+MLM+bbox)"),
    "large": dict(model="large", name="X2VLM-large (6 fusion layers)", batch=32, region_images=14, flop_pair=591.4e9,
                  flop_region=3 * 86.30e9, flop_image=3 * 123.11e9, what="pretrain step (ITC+ITM+MLM+bbox)"),
    "large12": dict(model="large12", name="X2VLM-large (12 fusion layers)", batch=32, region_images=14, flop_pair=738.3e9,
                    flop_region=3 * 147.45e9, flop_image=3 * 123.11e9, what="pretrain step (ITC+ITM+MLM+bbox)"),
    "video": dict(model="video", name="X2VLM-base video-text (8 frames)", batch=16, region_images=0, flop_pair=969.0e9,
                  flop_region=0.0, flop_image=0.0, what="video pretrain step (ITC+ITM+MLM), 8 frames avgpool"),
}


def model_config(kind):
    from x2vlm_b200 import pretrain
    if kind == "base":
        return pretrain.base_config()
    if kind == "large":    # configs/pretrain/x2vlm_large_1b_stage2.yaml: beit2-large + bert-large 12 text / 6 fusion layers
        return pretrain.large_config()
    if kind == "large12":  # BASELINE.json's literal "12 fusion layers"
        return pretrain.large_config(text_num_hidden_layers=24, text_fusion_start_at=12)
    if kind == "video":    # MSRVTT shape: 8 frames x 224 px, learned frame offsets, mean over time
        return pretrain.base_config(video_encoding="avgpool", frame_len=8, add_frame_pos=True)
    raise ValueError(kind)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


SETTLE_STEPS = 20  # untimed steps after the W warm-up steps (see run_ours)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_batches(batch, n_img, rank, pin, frames=0):
    from x2vlm_b200 import synth
    ib = synth.image_text_batch(batch, 40, seed=1234 + rank)
    if frames:
        ib["image"] = torch.randn(batch, frames, 3, 224, 224, generator=torch.Generator().manual_seed(99 + rank))
    rb = synth.region_batch(n_img, batch, 40, seed=4321 + rank) if n_img else None
    if pin:
        ib = {k: v.pin_memory() for k, v in ib.items()}
        rb = {k: v.pin_memory() for k, v in rb.items()} if rb is not None else None
    return ib, rb


def to_dev(d, dev):
    return {k: v.to(dev, non_blocking=True) for k, v in d.items()}


def nbytes(d):
    return sum(v.numel() * v.element_size() for v in d.values())


def run_ours(args):
    from x2vlm_b200 import accelerator, ops, pretrain
    from x2vlm_b200 import functional as XF
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))  # see X2kDDPAccelerator.set_up
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    C = CONFIGS[args.config]
    FLOP_IMAGE_PAIR, FLOP_REGION_SAMPLE, FLOP_VISION_IMAGE = C["flop_pair"], C["flop_region"], C["flop_image"]
    model = pretrain.XVLM(model_config(C["model"]))
    acc = accelerator.X2kDDPAccelerator({"lr": 1e-4, "weight_decay": 0.01, "bucket_mb": args.bucket_mb})
    ddp, opt, _ = acc.set_up(model, None, None, local, world, rank)
    ddp.train()
    XF.manual_seed(1234 + rank)
    B, n_img = args.batch, args.region_images
    ib_h, rb_h = make_batches(B, 0 if args.image_only else n_img, rank, pin=True, frames=8 if C["model"] == "video" else 0)

    def step(ib, rb, read_loss):
        with torch.no_grad():
            ddp.module.temp.clamp_(0.001, 0.5)  # Pretrain.py:327-328
        opt.zero_grad()
        losses = ddp.module.forward_mixed(ib, rb)
        loss = ddp.module.total_loss(losses)
        acc.backward_step(loss, opt)
        acc.optimizer_step(opt, ddp, 1.0)
        opt.step()
        return float(loss.detach()) if read_loss else loss

    host_ms = [0.0]

    graphed = None
    if not args.no_graph and not args.profile:
        def graph_body(inp):
            return step(inp["i"], inp.get("r"), False)
        ex = {"i": to_dev(ib_h, dev)}
        if rb_h is not None:
            ex["r"] = to_dev(rb_h, dev)
        try:
            graphed = acc.graph_step(graph_body, ex, optimizer=opt, warmup=3)
        except Exception as e:  # stay measurable: eager launches of the same kernels
            if rank == 0:
                sys.stderr.write("[bench] CUDA-graph capture failed, running eagerly: %r\n" % (e,))
            graphed = None
            torch.cuda.synchronize()
    eager_step = step
    if graphed is not None:
        def step(ib, rb, read_loss):  # noqa: F811  (same signature; inputs are copied into the graph's static buffers)
            loss = graphed({"i": ib, "r": rb} if rb is not None else {"i": ib})
            return float(loss.detach()) if read_loss else loss

    def timed(n, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        st.record()
        last = None
        t_host = time.perf_counter()
        host_batch = {"i": ib_h, "r": rb_h} if rb_h is not None else {"i": ib_h}
        if e2e and graphed is not None:
            graphed.stage(host_batch)  # step 0's inputs: this copy is exposed, the later ones overlap the previous replay
            loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
            loss_evt = [torch.cuda.Event() for _ in range(2)]
            pending = None
        for it in range(n):
            if e2e:
                if graphed is not None:
                    # every step: pinned host batch -> staging buffers (copy stream, prefetched during the previous
                    # replay like a data loader would) -> the graph's static buffers -> replay -> the step's loss copied
                    # to pinned host memory.  The host reads step k's loss after it has queued step k + 1 (one step of
                    # logging delay): a sync before the next launch would leave the GPU idle for the launch latency of an
                    # 8 000-node graph every step (measured: +4.4 ms / step on some boxes).
                    loss_t = graphed.run_staged()
                    loss_host[it & 1].copy_(loss_t.detach().reshape(()), non_blocking=True)
                    loss_evt[it & 1].record()
                    if it + 1 < n:
                        graphed.stage(host_batch)
                    if pending is not None:
                        loss_evt[pending].synchronize()
                        last = float(loss_host[pending])
                    pending = it & 1
                else:
                    last = step(to_dev(ib_h, dev), to_dev(rb_h, dev) if rb_h is not None else None, True)
            else:
                last = step(ib_d, rb_d, False)
        if e2e and graphed is not None and pending is not None:
            loss_evt[pending].synchronize()
            last = float(loss_host[pending])
        en.record()
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / n  # CPU time to enqueue a step (no sync inside unless e2e)
        torch.cuda.synchronize()
        ms = torch.tensor([st.elapsed_time(en)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launched = ops.launch_count() - l0
        if graphed is not None:  # replayed kernel nodes never pass through the library's host-side launch counter
            launched = graphed.x2k_launches_per_step * n
        return ms.item(), launched, float(last.detach()) if torch.is_tensor(last) else float(last)

    ib_d, rb_d = to_dev(ib_h, dev), (to_dev(rb_h, dev) if rb_h is not None else None)
    if args.profile:
        # one warm step, then exactly one step between cudaProfilerStart/Stop (ncu --profile-from-start off)
        step(ib_d, rb_d, False)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(ib_d, rb_d, False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for _ in range(max(args.warmup, 3)):
        step(ib_d, rb_d, False)
    # The pool's B200s run this step under a software power cap; the clock the cap settles at is reached about a second
    # into sustained load (the first timed region of a run was up to 4% slower than the same steps measured right after
    # it).  SETTLE_STEPS further untimed steps put the timed region into that steady state; they are reported in `config`.
    for _ in range(SETTLE_STEPS):
        step(ib_d, rb_d, False)
    torch.cuda.synchronize()
    if rank == 0:
        sys.stderr.write("[bench] warm-up done\n")
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, loss_v = timed(args.steps, e2e=False)
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop() if sampler else None
    sampler = ClockSampler(local) if rank == 0 else None
    ms_e2e, _, loss_e = timed(args.steps, e2e=True)
    clocks_e2e = sampler.stop() if sampler else None
    # the same resident-input steps once more, AFTER the end-to-end region: the pool's B200s run this step under a software
    # power cap, and the clock the cap settles at drifts over the first second of load — the difference between `value`
    # and `e2e` is only meaningful next to this number (reported in detail, never used as the headline)
    ms_again, _, _ = timed(args.steps, e2e=False)

    pairs_per_step = (B + (B if rb_h is not None else 0)) * world
    value = pairs_per_step * args.steps / (ms / 1e3)
    e2e_value = pairs_per_step * args.steps / (ms_e2e / 1e3)
    flop_step_gpu = B * FLOP_IMAGE_PAIR + ((B * FLOP_REGION_SAMPLE + n_img * FLOP_VISION_IMAGE) if rb_h is not None else 0.0)
    burst, sustained, hbm, src = peaks()
    step_tflops = flop_step_gpu * args.steps / (ms / 1e3) / 1e12  # per GPU (ms is the max over ranks)

    # image-only variant (64 pairs, 14.79 TFLOP / step): one more captured graph, N = 1 only
    image_only = None
    if world == 1 and rb_h is not None and graphed is not None and not args.skip_image_only and args.config == "base":
        try:
            g_img = acc.graph_step(lambda inp: eager_step(inp["i"], None, False), {"i": ib_d}, optimizer=opt, warmup=2)
            for _ in range(2):
                g_img({"i": ib_d})
            torch.cuda.synchronize()
            st_i, en_i = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st_i.record()
            for _ in range(args.steps):
                g_img({"i": ib_d})
            en_i.record()
            torch.cuda.synchronize()
            ms_i = st_i.elapsed_time(en_i) / args.steps
            image_only = {"ms_per_step": ms_i, "pairs_per_s": B / (ms_i / 1e3),
                          "step_tflops": B * FLOP_IMAGE_PAIR / (ms_i / 1e3) / 1e12}
            g_img.release()
        except Exception as e:
            sys.stderr.write("[bench] image-only variant skipped: %r\n" % (e,))

    roof = dominant_kernel_roofline(lambda: eager_step(ib_d, rb_d, False), sustained, src,  # every rank steps (collectives)
                                    table_path=args.kernel_table if rank == 0 else None)
    out = None
    graph_on = graphed is not None
    if rank == 0:
        sys.stderr.write("[bench] timed: %.2f ms/step resident (host enqueue %.2f ms/step), %.2f ms/step e2e\n"
                         % (ms / args.steps, host_enqueue_ms, ms_e2e / args.steps))
        base_cfg = args.config == "base"
        cpu = cpu_baseline(image_only=rb_h is None) if (world == 1 and not args.no_cpu_baseline and base_cfg) else None
        eager = None
        if world == 1 and not args.no_gpu_eager and base_cfg:
            try:
                if graphed is not None:
                    graphed.release()
                    graphed = None
                eager = gpu_eager_baseline(dev, B, n_img, rb_h is None)
            except Exception as e:  # the comparator must never cost the main line
                eager = {"unavailable": repr(e)[:300]}
        out = {
            "metric": metric_name(args), "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "detail": {"l2": "per-step working set (tens of GB of activations, 3 GB params/grads) >> 126 MB L2; no flush needed",
                       "loss_last_step": loss_v, "host_enqueue_ms_per_step": host_enqueue_ms,
                       "settle_steps_untimed": SETTLE_STEPS,
                       "cuda_graph": ("whole step (fwd+bwd+clip+AdamW) replayed as one CUDA graph; gpu_launches = x2k kernel "
                                      "nodes per replay x steps" if graph_on else "off (eager launches)"),
                       "algorithmic_tflop_per_step_per_gpu": flop_step_gpu / 1e12,
                       "step_tflops_per_gpu": step_tflops,
                       "step_frac_of_%s_sustained_bf16_peak" % src: step_tflops / sustained,
                       "image_iteration_only": image_only},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": nbytes(ib_h) + (nbytes(rb_h) if rb_h is not None else 0), "d2h_bytes_per_step": 4,
                    "loss_last_step": loss_e,
                    "sm_mhz": (clocks_e2e or {}).get("sm_mhz"), "power_w_max": (clocks_e2e or {}).get("power_w_max"),
                    "resident_ms_per_step_measured_after": ms_again / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        }
        if eager is not None:
            out["gpu_eager_baseline"] = eager
            if "value" in eager:
                out["gpu_eager_baseline"]["speedup_value_over_eager"] = value / eager["value"]
        if cpu is not None:
            out["cpu_baseline"] = cpu
    if rank == 0:
        emit_json(out)
    teardown(graphed, world)


def teardown(graphed, world):
    """Leave cleanly: NCCL's communicator destroy waits for every CUDA graph that captured one of its collectives to be
    released first (persistent references), so the step graph goes before the process group; a watchdog bounds it."""
    if graphed is not None:
        graphed.release()
    torch.cuda.synchronize()
    if world > 1:
        import threading
        done = threading.Event()

        def _destroy():
            try:
                dist.barrier()
                dist.destroy_process_group()
            finally:
                done.set()
        t = threading.Thread(target=_destroy, daemon=True)
        t.start()
        if not done.wait(60.0):
            sys.stderr.write("[bench] process-group teardown timed out; exiting\n")
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def dominant_kernel_roofline(step_fn, peak_sustained, src, table_path=None):
    """The dominant kernel is the tcgen05 GEMM (x2k gemm_tcgen05_pair_kernel / gemm_tcgen05_kernel: ~47% of the
    step's GPU time in profiles/).  One extra, instrumented step after the timed region brackets EVERY GEMM and
    attention launch with CUDA events on the launching stream; achieved = sum of algorithmic flops (2MNK per
    launch) / sum of launch durations = flops per launch / average launch duration.  The kernels run inside a long
    step, so the denominator is the sustained measured bf16 peak."""
    from x2vlm_b200 import ops
    with ops.kernel_timing() as kt:
        step_fn()
    tot = kt.totals()
    if table_path:  # per-shape table (markdown) for profiles/: which GEMM shapes / epilogues sit below the average
        rows = sorted(kt.by_shape().items(), key=lambda kv: -kv[1][2])
        with open(table_path, "w") as fh:
            fh.write("| family | shape / epilogue | launches | ms / step | avg us | TFLOP/s |\n|---|---|---:|---:|---:|---:|\n")
            for (fam, tag), (cnt, fl, t) in rows:
                fh.write("| %s | %s | %d | %.3f | %.1f | %.0f |\n" % (fam, tag, cnt, t, t / cnt * 1e3, fl / (t / 1e3) / 1e12))
    n, flops, ms = tot["gemm"]
    tf = flops / (ms / 1e3) / 1e12
    other = {k: {"launches": v[0], "ms_per_step": v[2], "tflops": v[1] / (v[2] / 1e3) / 1e12,
                 "frac": v[1] / (v[2] / 1e3) / 1e12 / peak_sustained} for k, v in tot.items() if k != "gemm"}
    return {"kernel": "x2k gemm_tcgen05_pair_kernel (+1-CTA gemm_tcgen05_kernel for M or N < 256), all %d launches of one step" % n,
            "bound": "tensor", "achieved": tf, "peak": peak_sustained, "unit": "TFLOP/s", "frac": tf / peak_sustained,
            "peak_source": src + " (sustained: kernels timed inside a long step)", "launches_per_step": n,
            "avg_launch_ms": ms / n, "ms_per_step": ms, "algorithmic_gflop_per_launch": flops / n / 1e9,
            "traffic": ROOFLINE_TRAFFIC, "other_kernels": other}


def metric_name(args):
    name = CONFIGS[getattr(args, "config", "base")]["name"].split(" (")[0].replace(" video-text", "")
    return "image-text pairs/sec pretrain step %s bf16" % name


def workload_config(args, world):
    """The static description of the workload — identical in both arms' JSON lines (`config`), so the driver's
    same-config check compares like with like; everything measured goes under `detail`."""
    C = CONFIGS[getattr(args, "config", "base")]
    B, n_img = args.batch, (0 if args.image_only else args.region_images)
    return {"workload": "%s %s, 224px, 40 tok, batch %d/GPU%s; random-init weights"
                        % (C["name"], C["what"], B, " image iteration only" if (args.image_only and C["region_images"]) else ""),
            "global_batch": B * world, "pairs_per_step_per_gpu": B if not n_img else 2 * B,
            "region_images_per_gpu": n_img, "seq_len": 40, "image_res": 224, "parallelism": "dp%d" % world,
            "optimizer": "AdamW(0.9, 0.98, eps 1e-8, wd 0.01) + clip 1.0"}


def reference_step_fn(device, batch_i, batch_r, n_img_r, autocast_bf16, threads=None):
    """One optimizer step of the UNMODIFIED reference with Pretrain.py:run_mixed_iter semantics (:189-252): the
    reference's own models.model_pretrain.XVLM on its own encoders (byte-code of /root/reference, oracle/_ref),
    its own optim.create_optimizer (optim.py:26-104), image iteration + region iteration summed, one backward,
    clip_grad_norm_ 1.0, AdamW; the per-loss .item() reads are the reference's metric_logger.update calls.
    Returns (step callable -> loss, pairs per step)."""
    import types
    from oracle import ref_shim
    from x2vlm_b200 import synth
    if threads:
        torch.set_num_threads(threads)
    model = ref_shim.build_reference_model("models.model_pretrain.XVLM", seed=0).to(device).train()
    import optim as ref_optim  # the reference's optim.py
    opt = ref_optim.create_optimizer(types.SimpleNamespace(lr=1e-4, weight_decay=0.01, lr_mult=2), model)
    ib = {k: v.to(device) for k, v in synth.image_text_batch(batch_i, 40, seed=1234).items()}
    rb = {k: v.to(device) for k, v in synth.region_batch(n_img_r, batch_r, 40, seed=4321).items()} if batch_r else None
    dev_type = torch.device(device).type

    def step():
        opt.zero_grad()
        with torch.no_grad():
            model.temp.clamp_(0.001, 0.5)  # Pretrain.py:327-328
        with torch.autocast(dev_type, dtype=torch.bfloat16, enabled=autocast_bf16):
            li = model(ib["image"], ib["text_ids"], ib["text_atts"], text_ids_masked=ib["text_ids_masked"],
                       masked_pos=ib["masked_pos"], masked_ids=ib["masked_ids"])
            loss = li["loss_itc"] + li["loss_itm"] + li["loss_mlm"]
            logged = [li[k].item() for k in ("loss_itc", "loss_itm", "loss_mlm")]
            if rb is not None:
                lr_ = model(rb["image"], rb["text_ids"], rb["text_atts"], text_ids_masked=rb["text_ids_masked"],
                            masked_pos=rb["masked_pos"], masked_ids=rb["masked_ids"], image_atts=rb["image_atts"],
                            idx_to_group_img=rb["idx_to_group_img"], target_bbox=rb["target_bbox"], is_image=rb["is_image"],
                            ret_bbox_loss=True)
                loss = loss + lr_["loss_itc"] + lr_["loss_itm"] + lr_["loss_mlm"] + lr_["loss_bbox"] + lr_["loss_giou"]
                logged += [lr_[k].item() for k in ("loss_itc", "loss_itm", "loss_mlm", "loss_bbox", "loss_giou")]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return sum(logged)

    return step, batch_i + batch_r


def gpu_eager_baseline(dev, batch, n_img, image_only, steps=5, warmup=3):
    """The >= 6x comparator of BASELINE.md §4.4: the reference's own modules, PyTorch eager under
    torch.autocast('cuda', bfloat16) (apex O1 is not installable offline; bf16 needs no loss scaler), same GPU, same
    batch (64 + 64 over 26 images), inputs resident on the device, CUDA-event timed after warm-up."""
    from oracle import ref_shim
    if not ref_shim.available():
        return {"unavailable": "oracle/_ref (compiled reference) not present"}
    step, pairs = reference_step_fn(dev, batch, 0 if image_only else batch, n_img, autocast_bf16=True)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(steps):
        loss = step()
    en.record()
    torch.cuda.synchronize()
    ms = st.elapsed_time(en) / steps
    return {"value": pairs / (ms / 1e3), "unit": "pairs/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "what": "UNMODIFIED reference (models.model_pretrain.XVLM on models.beit2/models.xbert, optim.create_optimizer), "
                    "PyTorch eager, torch.autocast(cuda, bfloat16), run_mixed_iter semantics, batch %d%s, resident inputs, "
                    "same GPU as `value`" % (batch, "" if image_only else " + %d regions / %d images" % (batch, n_img)),
            "loss_last_step": loss, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}


def host_threads():
    """CPU threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def oracle_step_fn(batch_i, batch_r, threads):
    """fwd + bwd + AdamW of the reference algorithm (oracle/restate.py, fp32, torch CPU) on a small mixed batch."""
    from oracle import restate
    from x2vlm_b200 import pretrain, synth
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = pretrain.XVLM(pretrain.base_config())  # parameter container only (CPU, never forwarded)
    sd = {k: v for k, v in m.state_dict().items()}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full = dict(sd); full.update(leaves)
    full["text_encoder.cls.predictions.decoder.weight"] = full["text_encoder.bert.embeddings.word_embeddings.weight"]
    full["text_encoder.cls.predictions.decoder.bias"] = full["text_encoder.cls.predictions.bias"]
    params = list({id(v): v for v in leaves.values()}.values())
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    ib = synth.image_text_batch(batch_i, 40, seed=1234)
    rb = synth.region_batch(max(1, batch_r // 2), batch_r, 40, seed=4321) if batch_r else None
    shp = restate.Shapes()

    def step():
        opt.zero_grad()
        li = restate.pretrain_forward(full, shp, ib["image"], ib["text_ids"], ib["text_atts"], ib["text_ids_masked"],
                                      ib["masked_pos"], ib["masked_ids"], *synth.hard_negative_indices(batch_i, 1), train=True)
        loss = sum(li.values())
        if rb is not None:
            lr_ = restate.pretrain_forward(full, shp, rb["image"], rb["text_ids"], rb["text_atts"], rb["text_ids_masked"],
                                           rb["masked_pos"], rb["masked_ids"], *synth.hard_negative_indices(batch_r, 2),
                                           image_atts=rb["image_atts"], idx_to_group_img=rb["idx_to_group_img"],
                                           target_bbox=rb["target_bbox"], is_image=rb["is_image"], ret_bbox_loss=True, train=True)
            loss = loss + sum(lr_.values())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return float(loss)

    return step, batch_i + batch_r


CPU_SAMPLE_I, CPU_SAMPLE_R, CPU_SAMPLE_IMGS = 8, 8, 3  # bounded sample of the 64 + 64 / 26-image step (same 26/64 image ratio)


def cpu_step_fn(image_only):
    """(step, pairs per step, kind, cores, sample text) of the CPU arm: the UNMODIFIED reference (oracle/_ref byte-code of
    /root/reference, fp32, all host threads — BASELINE.md §4.3: B = 8 image iteration, here together with the region
    iteration of the mixed step) or, only if that copy is missing, the oracle port."""
    from oracle import ref_shim
    threads = host_threads()
    bi, br = CPU_SAMPLE_I, (0 if image_only else CPU_SAMPLE_R)
    if ref_shim.available():
        step, pairs = reference_step_fn("cpu", bi, br, CPU_SAMPLE_IMGS, autocast_bf16=False, threads=threads)
        kind = "reference"
        what = "UNMODIFIED reference (models.model_pretrain.XVLM + optim.create_optimizer, run_mixed_iter semantics)"
    else:
        step, pairs = oracle_step_fn(bi, br, threads)
        kind, what = "port", "oracle/restate.py port (oracle/_ref missing)"
    sample = ("each step = one optimizer step on %d image-text pairs%s — a bounded sample of the batch-64 workload; %s, fp32, "
              "torch CPU, %d threads" % (bi, "" if image_only else " + %d region samples over %d images" % (br, CPU_SAMPLE_IMGS),
                                         what, threads))
    return step, pairs, kind, threads, sample


def cpu_baseline(budget_s=25.0, image_only=False):
    """Bounded sample: one warm-up step, then timed steps until ~budget_s of CPU work (at least one)."""
    step, pairs, kind, threads, sample = cpu_step_fn(image_only)
    step()  # warm-up (the first step pays allocator / lazy-init costs)
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 8:
            break
    return {"value": pairs * n / dt, "unit": "pairs/s", "cores": threads, "kind": kind,
            "sample": "%d timed step(s) in %.1f s after 1 warm-up; %s" % (n, dt, sample)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores — the unmodified
    reference itself (byte-code compiled from /root/reference by oracle/build_ref.py; it travels to the GPU box in
    oracle/_ref/), same metric / unit / config as the main arm, every step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, pairs, kind, threads, sample = cpu_step_fn(args.image_only)
    warm = max(args.warmup, 0)
    for _ in range(warm):
        step()
    n = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(n):
        loss = step()
    dt = time.perf_counter() - t0
    v = pairs * n / dt
    emit_json({
        "impl": "reference", "metric": "image-text pairs/sec pretrain step X2VLM-base bf16", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": n, "warmup": warm, "ms_per_step": dt / n * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "detail": {"loss_last_step": loss, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def run_retrieval(args):
    """--config retrieval (BASELINE config 4): 1 000 images x 5 000 captions, ITC all-pairs + top-128 ITM re-rank in both
    directions (Retrieval.py:71-157) on one GPU.  A "step" is one whole evaluation; value = re-ranked (image, caption)
    pairs per second.  Reference-algorithm FLOPs: 35.1 T vision + 34.3 T text + 768 000 x 6.93 G rerank."""
    from x2vlm_b200 import ops, pretrain, retrieval, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_img, n_txt, k = args.retrieval_images, args.retrieval_texts, args.k_test
    torch.manual_seed(0)
    model = pretrain.XVLM(pretrain.base_config()).to(dev).eval()
    b = synth.image_text_batch(n_txt, 40, seed=1234)
    g = torch.Generator().manual_seed(5)
    images = torch.randn(n_img, 3, 224, 224, generator=g).pin_memory()
    ids, atts = b["text_ids"].pin_memory(), b["text_atts"].pin_memory()

    def one(resident):
        im, ti, ta = (images.to(dev, non_blocking=True), ids.to(dev, non_blocking=True), atts.to(dev, non_blocking=True)) \
            if not resident else resident
        s_i2t, s_t2i, sims = retrieval.evaluation(model, im, ti, ta, k_test=k, image_bs=100, text_bs=500,
                                                  rows_per_call=args.rows_per_call)
        return s_i2t, s_t2i

    res = (images.to(dev), ids.to(dev), atts.to(dev))
    for _ in range(max(1, min(args.warmup, 1))):
        one(res)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    l0 = ops.launch_count()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(1, args.steps)
    st.record()
    for _ in range(n):
        s_i2t, s_t2i = one(res)
    en.record()
    torch.cuda.synchronize()
    ms = st.elapsed_time(en) / n
    launches = ops.launch_count() - l0
    clocks = sampler.stop()
    st.record()
    for _ in range(n):   # end to end: pinned host images / token ids -> device, both score matrices back on the host
        s_i2t, s_t2i = one(None)
        h1, h2 = s_i2t.cpu(), s_t2i.cpu()
    en.record()
    torch.cuda.synchronize()
    ms_e2e = st.elapsed_time(en) / n
    pairs = n_img * min(k, n_txt) + n_txt * min(k, n_img)
    flop_ref = n_img * 35.13e9 + n_txt * 6.85e9 + pairs * 6.93e9            # reference algorithm (K/V re-projected per pair)
    flop_exec = n_img * 35.13e9 + n_txt * 6.85e9 + pairs * (6.93e9 - 2.79e9) + n_img * 2.79e9  # i2t: K/V once per image
    _, sustained, _, src = peaks()
    out = {"metric": "image-text pairs/sec retrieval rerank X2VLM-base bf16", "value": pairs / (ms / 1e3), "unit": "pairs/s",
           "n_gpus": 1, "steps": n, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "retrieval inference: %d x 224px images vs %d 40-token captions, ITC all-pairs + top-%d ITM "
                                  "cross-attention rerank both ways; random-init weights" % (n_img, n_txt, k),
                      "reranked_pairs": pairs, "rows_per_call": args.rows_per_call},
           "detail": {"reference_algorithm_tflop": flop_ref / 1e12, "executed_tflop_lower_bound": flop_exec / 1e12,
                      "redundant_kv_tflop_removed_i2t": (n_img * min(k, n_txt) - n_img) * 2.79e9 / 1e12,
                      "tflops_reference_algorithm": flop_ref / (ms / 1e3) / 1e12,
                      "frac_of_%s_sustained_bf16_peak" % src: flop_ref / (ms / 1e3) / 1e12 / sustained,
                      "note": "t2i candidates are de-duplicated per call: each distinct image's K/V is projected once per call"},
           "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": images.numel() * 4 + ids.numel() * 8 + atts.numel() * 8,
                   "d2h_bytes_per_step": 2 * n_img * n_txt * 4},
           "gpu_launches": launches, "clocks": clocks}
    if args.retrieval_eager:
        try:
            from oracle import ref_shim
            ref = ref_shim.build_reference_model().to(dev).eval()
            # the loop writes ITM scores into an fp32 matrix (Retrieval.py:133): hand it fp32 scores under autocast
            ref.itm_head.register_forward_hook(lambda mod, inp, out: out.float())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ref_shim.run_reference_retrieval(ref, res[0], res[1], res[2], k, "cuda", image_bs=100, text_bs=500)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out["gpu_eager_baseline"] = {"value": pairs / dt, "unit": "pairs/s", "ms_per_step": dt * 1e3,
                                         "what": "UNMODIFIED reference Retrieval.evaluation loop on the reference's modules, "
                                                 "PyTorch eager, torch.autocast(cuda, bfloat16), same GPU, resident inputs",
                                         "speedup_value_over_eager": (pairs / (ms / 1e3)) / (pairs / dt)}
        except Exception as e:
            out["gpu_eager_baseline"] = {"unavailable": repr(e)[:300]}
    emit_json(out)


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG is set, as it is on the GPU boxes), so file descriptor 1 is pointed at stderr for the whole run and the
    JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="base", choices=sorted(CONFIGS) + ["retrieval"],
                    help="BASELINE.json workload: base (headline), large / large12, video, retrieval (1k x 5k, k = 128)")
    ap.add_argument("--retrieval-images", type=int, default=1000)
    ap.add_argument("--retrieval-texts", type=int, default=5000)
    ap.add_argument("--k-test", type=int, default=128)
    ap.add_argument("--rows-per-call", type=int, default=8)
    ap.add_argument("--retrieval-eager", action="store_true", help="also time the reference's own evaluation loop (eager, bf16)")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--region-images", type=int, default=None)
    ap.add_argument("--image-only", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=48.0)
    ap.add_argument("--nccl-max-ctas", type=int, default=8, help="NCCL_MAX_CTAS for the gradient all-reduce (N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-PyTorch-eager comparator (N = 1 leg)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the step graph")
    ap.add_argument("--skip-image-only", action="store_true", help="do not also time the image-iteration-only step (N = 1)")
    ap.add_argument("--kernel-table", default=None, help="write the per-shape GEMM / attention timing table (markdown) here")
    ap.add_argument("--profile", action="store_true", help="one step inside cudaProfilerStart/Stop, no JSON (for ncu)")
    args = ap.parse_args()
    if args.config == "retrieval":
        return run_retrieval(args)
    if args.batch is None:
        args.batch = CONFIGS[args.config]["batch"]
    if args.region_images is None:
        args.region_images = CONFIGS[args.config]["region_images"]
    if not args.region_images:
        args.image_only = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
