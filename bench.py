#!/usr/bin/env python
"""Benchmark of the OCTAve scribble-supervised training step (BASELINE.json metric: train images/s at 400x400).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A step = one adversarial iteration on a synthetic OCTA batch: G-step (segmentor forward, fused WPCE + KLD + LS-G
through the mask critic, backward, gradient all-reduce, SGD) followed by the D-step (LS-D on real/fake pyramids,
backward, SGD).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (batch per GPU, H, W)
    "c1": (2, 304, 304),
    "c2": (32, 400, 400),
    "c4": (64, 304, 304),
    "c5": (8, 1024, 1024),
}
FLOP_PER_IMG_FWD = {304: 91.42e9 + 0.83e9 * 3, 400: 158.00e9 + 1.43e9 * 3, 1024: 1030.09e9 + 9.44e9 * 3}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason sampling during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port: torch fp32 functional restatement, oracle/octave_oracle.py)
# ----------------------------------------------------------------------------------------------------
def cpu_step_fn(H: int, W: int, batch: int, seed: int = 0):
    from octave_b200 import synth
    from oracle import octave_oracle as O
    from octave_b200.model import OctaScribbleNet
    torch.manual_seed(seed)
    net = OctaScribbleNet(torch.Size((batch, 3, H, W)), torch.Size((batch, 2, H, W)), True, False, instance_noise=False, label_noise=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    del net
    seg = {k[len("segmentor."):]: (v.requires_grad_() if v.is_floating_point() and "running" not in k else v)
           for k, v in sd.items() if k.startswith("segmentor.")}
    dis = {k[len("discriminator."):]: (v.requires_grad_() if v.is_floating_point() and not k.endswith(("_u", "_v")) else v)
           for k, v in sd.items() if k.startswith("discriminator.")}
    x, ys, _ = synth.octa_batch(batch, H, W, seed=seed)
    real = synth.mask_pyramid(batch, H, W)
    seg_p = [v for k, v in seg.items() if v.requires_grad and not k.startswith("linear_head_")]
    dis_p = [v for v in dis.values() if v.requires_grad]

    def step():
        att, agg, _ = O.segmentor_forward(seg, x, training=True)
        lg = O.weighted_partial_ce(torch.softmax(agg, 1), ys, 2) + 0.1 * O.interlayer_divergence(att) + \
            0.1 * O.ls_generator_loss(O.discriminator_forward(dis, att, depth=4, training=True))
        gs = torch.autograd.grad(lg, seg_p, allow_unused=True)
        ld = O.ls_discriminator_loss(O.discriminator_forward(dis, real, depth=4, training=True),
                                     O.discriminator_forward(dis, [a.detach() for a in att], depth=4, training=True))
        gd = torch.autograd.grad(ld, dis_p, allow_unused=True)
        with torch.no_grad():
            for p, g in zip(seg_p, gs):
                if g is not None:
                    p.add_(g, alpha=-1e-3)
            for p, g in zip(dis_p, gd):
                if g is not None:
                    p.add_(g, alpha=-1e-3)
        return float(lg)

    return step


def run_torch_gpu(H, W, batch, steps, warmup, autocast: bool):
    """On-box comparator: the SAME reference arithmetic (oracle port of the reference modules, stock
    torch ops: cuDNN convolutions, ATen BatchNorm / softmax) on the GPU.  Baseline leg only — never the product path."""
    from octave_b200 import synth
    from oracle import octave_oracle as O
    from octave_b200.model import OctaScribbleNet
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = OctaScribbleNet(torch.Size((batch, 3, H, W)), torch.Size((batch, 2, H, W)), True, False, instance_noise=False, label_noise=False)
    sd = {k: v.detach().clone().to(dev) for k, v in net.state_dict().items()}
    del net
    seg = {k[len("segmentor."):]: (v.requires_grad_() if v.is_floating_point() and "running" not in k else v)
           for k, v in sd.items() if k.startswith("segmentor.")}
    dis = {k[len("discriminator."):]: (v.requires_grad_() if v.is_floating_point() and not k.endswith(("_u", "_v")) else v)
           for k, v in sd.items() if k.startswith("discriminator.")}
    x, ys, _ = synth.octa_batch(batch, H, W, seed=0, n_ridges=8)
    real = [r.to(dev) for r in synth.mask_pyramid(batch, H, W, n_ridges=8)]
    x, ys = x.to(dev).contiguous(memory_format=torch.channels_last), ys.to(dev)
    seg_p = [v for k, v in seg.items() if v.requires_grad and not k.startswith("linear_head_")]
    dis_p = [v for v in dis.values() if v.requires_grad]

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            att, agg, _ = O.segmentor_forward(seg, x, training=True)
        att = [a.float() for a in att]
        lg = O.weighted_partial_ce(torch.softmax(agg.float(), 1), ys, 2) + 0.1 * O.interlayer_divergence(att) + \
            0.1 * O.ls_generator_loss(O.discriminator_forward(dis, att, depth=4, training=True))
        gs = torch.autograd.grad(lg, seg_p, allow_unused=True)
        ld = O.ls_discriminator_loss(O.discriminator_forward(dis, real, depth=4, training=True),
                                     O.discriminator_forward(dis, [a.detach() for a in att], depth=4, training=True))
        gd = torch.autograd.grad(ld, dis_p, allow_unused=True)
        with torch.no_grad():
            torch._foreach_add_([p for p, g in zip(seg_p, gs) if g is not None], [g for g in gs if g is not None], alpha=-1e-3)
            torch._foreach_add_([p for p, g in zip(dis_p, gd) if g is not None], [g for g in gd if g is not None], alpha=-1e-3)
        return lg

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "batch": batch,
            "what": "oracle port of the reference modules on stock torch CUDA ops (cuDNN), " + ("bf16 autocast" if autocast else "fp32 (TF32 off)")
                    + ", same G+D step; comparator only"}


def run_cpu(H, W, sample_batch, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(H, W, sample_batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return {"value": sample_batch / sec, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"batch {sample_batch} of the {H}x{W} workload per step, fp32, median of {steps} after {warmup} warm-up; "
                      f"oracle port of the reference modules (/root/reference is not on the GPU box)", "s_per_step": sec}


# ----------------------------------------------------------------------------------------------------
def kernel_rooflines(peaks, B, H, W):
    """Times the dominant tensor-core conv and the fused loss kernel in isolation with CUDA events on the launch stream."""
    import ctypes as C
    from octave_b200 import _lib, ops, losses
    from octave_b200.ops import Act, ConvSpec
    dev = torch.device("cuda")
    out = {}
    # (1) decoder_2.conv.0: 3x3 512->256 at H/4 (23.6 GFLOP/img at 400^2 — one of the three fat decoder convs)
    h, w = H // 4, W // 4
    x = Act(torch.randn(B, h, w, 512, device=dev).bfloat16(), B, h, w, 512)
    wt = torch.nn.Parameter(torch.randn(256, 512, 3, 3, device=dev) * 0.02)
    spec = ConvSpec(wt, None, 512, 256, 3, 1, 1, 1)
    y = Act.empty(B, h, w, 256, torch.bfloat16, dev)
    flops = 2.0 * B * h * w * 256 * 512 * 9
    for _ in range(3):
        ops.conv_fwd(x, spec, out=y)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        ops.conv_fwd(x, spec, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ach = flops / (ms * 1e-3) / 1e12
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this launch at B=32, 400^2 from one `ncu --set full` capture
    # (profiles/ncu_conv_fat_512x256_bn256_r01.txt: 330.4 MB + 137.7 MB; algorithmic minimum x + y + w = 494 MB)
    traffic = 468.1e6 if (B, H, W) == (32, 400, 400) else None
    out["roofline"] = {"kernel": "conv_tc_kernel<256,64,4> (decoder_2.conv.0 fwd 3x3 512->256)", "bound": "tensor",
                       "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                       "traffic": traffic, "algorithmic_flop": flops,
                       "peak_source": peaks["source"] + " burst (kernel timed alone)", "ms_per_launch": ms}
    # (2) fused loss kernel, forward statistics + gradient pass, bf16 maps (36.0 B/pixel algorithmic)
    g = torch.Generator(device=dev).manual_seed(0)
    agg = torch.randn(B, 2, H, W, device=dev, generator=g).bfloat16().requires_grad_()
    ys = (torch.rand(B, 2, H, W, device=dev, generator=g) < 0.03).to(torch.bfloat16)
    att = [torch.softmax(torch.randn(B, 2, H >> k, W >> k, device=dev, generator=g), 1).bfloat16().requires_grad_() for k in range(5)]
    fl = losses.FusedSegmentorLoss()
    def once():
        r = fl(agg, ys, att)
        (r['supervised'] + r['divergence']).backward()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    # time the two launches themselves through the C-ABI (no autograd / allocator noise)
    cfg = losses._LossCfg(_lib.LOSS_WPCE | _lib.LOSS_KLD | _lib.LOSS_FROM_LOGITS, att_weights=[1.0] * 4, sum_weights=4.0)
    attd = [a.detach() for a in att]
    desc = losses._build_desc(cfg, agg.detach(), attd, None, None)
    stats = torch.empty(_lib.lib.octave_loss_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
    outv = torch.empty(8, device=dev)
    gs = torch.ones(8, device=dev)
    g_y = torch.empty_like(agg); g_a = [torch.empty_like(a) for a in attd]
    arr, garr = losses._ptr_array(attd), losses._ptr_array(g_a)
    sp = torch.cuda.current_stream().cuda_stream
    def launch():
        _lib.lib.octave_loss_fwd(C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, None, stats.data_ptr(), outv.data_ptr(), sp)
        _lib.lib.octave_loss_bwd(C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, None, stats.data_ptr(), gs.data_ptr(),
                                 g_y.data_ptr(), garr, None, None, sp)
    for _ in range(3):
        launch()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()          # L2 flush between timed iterations (B200 L2 = 126 MB)
        torch.cuda.synchronize(); e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    nbytes = 17.99 * 2 * B * H * W
    gbs = nbytes / (ms * 1e-3) / 1e9
    out["loss_kernel"] = {"kernel": "loss_fast_fwd_kernel<bf16> + loss_fast_bwd_kernel<bf16>", "bound": "hbm", "achieved": gbs,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                          "algorithmic_bytes": nbytes, "ms_fwd_plus_bwd": ms, "peak_source": peaks["source"]}
    return out


def run_gpu(args):
    import torch.distributed as dist
    from octave_b200 import _lib, config, synth
    from octave_b200.model import OctaScribbleNet
    from octave_b200.train import TrainStep
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # control plane (barriers, max over ranks of the timings) on gloo: the NCCL communicator only ever carries the
        # gradient all-reduces, which in graph mode are replayed from the captured step
        ctl = dist.new_group(backend="gloo")
    B, H, W = CONFIGS[args.config]
    if args.batch:
        B = args.batch
    config.set_compute_dtype(args.dtype)
    config.nan_check = False
    torch.manual_seed(0)
    net = OctaScribbleNet(torch.Size((B, 3, H, W)), torch.Size((B, 2, H, W)), True, False).to(dev).train()
    ts = TrainStep(net, distributed=world > 1)
    # synthetic batches in pinned host memory (different data per rank = weak scaling)
    nb = 2
    host = []
    for i in range(nb):
        x, ys, _ = synth.octa_batch(B, H, W, seed=100 * rank + i, n_ridges=12)
        real = synth.mask_pyramid(B, H, W, seed=1000 + 100 * rank + i, n_ridges=12)
        host.append((x.pin_memory(), ys.pin_memory(), [r.pin_memory() for r in real]))
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 4 + sum(r.numel() * 4 for r in host[0][2])

    # one CUDA graph per step, also data-parallel: the bucketed NCCL all-reduces are captured with the backward pass
    use_graph = not args.no_graph
    from octave_b200.train import HostPrefetcher
    pre = HostPrefetcher(dev)

    def one_step(i, e2e: bool, dev_batches=None):
        if e2e:
            x, ys, real = host[i % nb]          # pinned host buffers: the H2D copies are part of the step
            if not use_graph:
                got = pre.get()                 # uploaded by the copy stream while the previous step computed
                if got is None:
                    pre.put(x, ys, real)
                    got = pre.get()
                x, ys, real = got
        else:
            x, ys, real = dev_batches[i % nb]
        res = ts.step_graphed(x, ys, real) if use_graph else ts.step(x, ys, real)
        if e2e:                                 # input pipelining: next batch's H2D overlaps this step's compute
            if use_graph:
                ts.prefetch(*host[(i + 1) % nb])
            else:
                pre.put(*host[(i + 1) % nb])
        if e2e:
            return float(res['total'].item())     # device -> host read of the step's loss
        return res['total']

    def timed(e2e: bool, steps: int, warmup: int):
        dev_batches = None
        if not e2e:
            dev_batches = [(x.to(dev), ys.to(dev), [r.to(dev) for r in rr]) for x, ys, rr in host]
        for i in range(warmup):
            one_step(i, e2e, dev_batches)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=ctl)
        torch.cuda.synchronize()
        l0 = _lib.lib.octave_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one_step(i, e2e, dev_batches)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=ctl)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)])
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=ctl)
        n_launch = _lib.lib.octave_launch_count() - l0
        if use_graph and getattr(ts, "_graph", None) is not None:
            n_launch = steps * ts.graph_kernel_nodes      # replayed kernel nodes of this library (host-side counter sees none)
        return float(ms.item()) / steps, n_launch

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(False, args.steps, args.warmup)
    ms_e2e, _ = timed(True, args.steps, max(1, args.warmup // 2))
    if sampler:
        sampler.stop_flag.set(); sampler.join(timeout=3)
    peaks = load_peaks()
    line = {
        "metric": f"train images/s, {H}x{W} synthetic OCTA (adversarial step: G-step + D-step)",
        "value": world * B / (ms_step * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{args.config}: OctaScribbleNet G+D training step, batch {B}/GPU, {H}x{W}, random-init weights",
                   "global_batch": world * B, "parallelism": f"dp{world}", "l2": "working set >> L2 (inputs+activations of one step are GBs)",
                   "optimizer": "torch SGD(momentum), foreach; the per-step bf16 operand re-pack of all conv weights is one multi-tensor kernel launch (SURVEY.md §8f.1)"},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
    }
    line["config"]["cuda_graph"] = bool(use_graph and getattr(ts, "_graph", None) is not None)
    if use_graph and getattr(ts, "graph_error", None):
        line["config"]["cuda_graph_error"] = ts.graph_error
    if rank == 0:
        line["clocks"] = sampler.summary()
        flops = 3.0 * FLOP_PER_IMG_FWD.get(H, 158e9 * (H * W) / 160000.0) * B
        line["step_tensor_fraction"] = {"algorithmic_tflop_per_step": flops / 1e12,
                                        "achieved_tflops": flops / (ms_step * 1e-3) / 1e12,
                                        "frac_of_sustained_peak": flops / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]}
        if args.dtype == "bf16":
            try:
                line.update(kernel_rooflines(peaks, min(B, 32), H, W))
            except Exception as e:  # pragma: no cover
                line["roofline_error"] = repr(e)
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = run_cpu(H, W, 2, 3, 1)
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"error": repr(e)}
        if world == 1 and args.comparator:
            del ts, net
            torch.cuda.empty_cache()
            for name, ac in (("torch_gpu_bf16_autocast", True), ("torch_gpu_fp32", False)):
                try:
                    line[name] = run_torch_gpu(H, W, B, 3, 2, ac)
                except Exception as e:  # pragma: no cover
                    line[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        print(json.dumps(line))
    if world > 1:
        sys.stdout.flush()
        dist.barrier(group=ctl)
        if use_graph:
            os._exit(0)          # a communicator whose collectives live in a captured graph is not torn down here
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    B, H, W = CONFIGS[args.config]
    cb = run_cpu(H, W, 2, max(1, min(args.steps, 5)), max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": f"train images/s, {H}x{W} synthetic OCTA (adversarial step: G-step + D-step)",
            "value": cb["value"], "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{args.config}: OctaScribbleNet G+D training step, {H}x{W}; each step is a bounded sample of batch 2"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--comparator", action="store_true", help="also time the reference arithmetic on stock torch CUDA ops (cuDNN) on this GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch the ~1100 kernels of a step eagerly instead of replaying one CUDA graph")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
