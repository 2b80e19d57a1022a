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
    """One G+D step of the reference algorithm on the host: oracle port (oracle/octave_oracle.py) over a seeded state dict
    built from the committed shape table (oracle/synth_state.py).  Nothing of octave_b200's CUDA library is imported here
    (octave_b200.synth is plain torch)."""
    from octave_b200 import synth
    from oracle import octave_oracle as O
    from oracle import synth_state
    sd = synth_state.seeded_state(H, W, seed)
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
RIDGE_FLOP_PER_BYTE = 218.0     # sustained bf16 peak / measured HBM bandwidth: above = tensor-bound


def family_rooflines(ts, batch, peaks, step_ms):
    """One extra EAGER step with CUDA events around every C-ABI call (octave_b200.profiler) on the launching stream, grouped
    into kernel families.  Per family: algorithmic FLOPs or bytes of its launches / their summed duration, against the
    sustained tensor peak (the kernels run inside a long step) or the measured HBM bandwidth.  Algorithmic units (DESIGN.md
    §3): a convolution = 2*pixels*Cout*Cin/g*k^2 FLOP and 2 B * pixels * (Cin + Cout) of minimum traffic; an elementwise
    BatchNorm pass = element size * B*H*W*C per tensor it must read or write."""
    from octave_b200 import profiler
    x, ys, real = batch

    def eager():
        feed = getattr(ts, "_feed", None)      # graph mode installed a host random feed on the critic: fill it as a replay would
        if feed is not None:
            feed.draw(); feed.upload()
        ts.step(x, ys, real)

    # this profile runs on rank 0 alone: no collective may be entered (the other ranks are not stepping); and every kernel
    # runs alone on one stream (config.overlap_wgrad off), so that an event pair brackets exactly one kernel's execution
    from octave_b200 import config as _cfg
    reducer, hook, ov = ts.reducer, ts.net.segmentor._grad_ready_hook, _cfg.overlap_wgrad
    ts.reducer, ts.net.segmentor._grad_ready_hook, _cfg.overlap_wgrad = None, None, False
    profiler.enable()
    try:
        eager()
        profiler.reset()
        eager()
        recs = profiler.records()
    finally:
        profiler.disable()
        ts.reducer, ts.net.segmentor._grad_ready_hook, _cfg.overlap_wgrad = reducer, hook, ov
    fam = {}

    def add(name, bound, ms, flop=0.0, byts=0.0):
        f = fam.setdefault(name, {"bound": bound, "ms": 0.0, "flop": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += ms; f["flop"] += flop; f["bytes"] += byts; f["launches"] += 1

    total_ms = 0.0
    for key, ms, d in recs:
        total_ms += ms
        n = d["name"]
        if d.get("kind") == "conv":
            cout = d["cout"] * (4 if d["mode"] == 1 else 1)
            pix = d["B"] * d["H"] * d["W"]
            flop = 2.0 * pix * cout * (d["cin"] // d["g"]) * d["k"] ** 2
            byts = 2.0 * pix * (d["cin"] + cout)
            tensor = flop / byts >= RIDGE_FLOP_PER_BYTE
            what = "weight gradient" if n.endswith("wgrad") else "forward / data gradient"
            add(f"conv {what}, {'tensor' if tensor else 'hbm'}-bound shapes ({'conv_tc_wgrad_kernel / conv3x3_halo_wgrad_kernel' if n.endswith('wgrad') else 'conv_tc_kernel / conv3x3_halo_kernel'})",
                "tensor" if tensor else "hbm", ms, flop, byts)
        elif d.get("kind") == "act" and n in ("octave_affine_act", "octave_bn_bwd_reduce", "octave_bn_bwd_apply"):
            el = d["esize"] * d["B"] * d["H"] * d["W"] * d["C"]
            pres = d["present"]
            if n == "octave_affine_act":          # x, res?, y
                passes = 2 + (1 if 2 in pres else 0)
            elif n == "octave_bn_bwd_reduce":     # dy, mask?, x, dmasked?
                passes = 2 + (1 if 1 in pres else 0) + (1 if 6 in pres else 0)
            else:                                 # dy, mask?, x, dx, dmasked?
                passes = 3 + (1 if 1 in pres else 0) + (1 if 11 in pres else 0)
            add({"octave_affine_act": "BatchNorm apply (+residual, +ReLU) (affine_act_kernel)",
                 "octave_bn_bwd_reduce": "BatchNorm backward, reduction pass (bn_bwd_reduce_kernel)",
                 "octave_bn_bwd_apply": "BatchNorm backward, apply pass (bn_bwd_apply_kernel)"}[n], "hbm", ms, 0.0, float(passes * el))
        else:
            add("other kernels of this library (split-attention, pools, heads, critic plumbing, losses, packing)", "-", ms)
    out = []
    for name, f in fam.items():
        rec = {"family": name, "bound": f["bound"], "launches": f["launches"], "ms_per_step": f["ms"], "share_of_kernel_time": f["ms"] / total_ms}
        if f["bound"] == "tensor":
            ach = f["flop"] / (f["ms"] * 1e-3) / 1e12
            rec.update(achieved=ach, peak=peaks["bf16_tflops_sustained"], unit="TFLOP/s", frac=ach / peaks["bf16_tflops_sustained"],
                       algorithmic_flop=f["flop"])
        elif f["bound"] == "hbm":
            ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
            rec.update(achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"], algorithmic_bytes=f["bytes"])
        out.append(rec)
    out.sort(key=lambda r: -r["ms_per_step"])
    return out, total_ms


def best_conv_roofline(peaks, B, H, W):
    """The fat decoder convolution (decoder_2.conv.0: 3x3 512->256 at H/4, 23.6 GFLOP/img at 400^2) timed alone."""
    from octave_b200 import ops
    from octave_b200.ops import Act, ConvSpec
    dev = torch.device("cuda")
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
    # traffic of this launch at B=32, 400^2: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture
    # (profiles/ncu_conv_fat_512x256_bn256_r01.txt: 330.4 MB + 137.7 MB; algorithmic minimum x + y + w = 494 MB)
    traffic = 468.1e6 if (B, H, W) == (32, 400, 400) else None
    return {"kernel": "conv_tc_kernel<256,64,4> (decoder_2.conv.0 fwd 3x3 512->256), timed alone", "bound": "tensor", "achieved": ach,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"], "traffic": traffic,
            "algorithmic_flop": flops, "peak_source": peaks["source"] + " burst (kernel timed alone)", "ms_per_launch": ms}


def loss_rooflines(peaks, B, H, W):
    """K9, the G-step loss (WPCE + KLD on logits, 5-level pyramid): ONE co-resident launch (label counts, then the values-and-
    gradients sweep, then the finalisation by the last block), through the C-ABI.  Rotating input sets larger than L2, launches queued back to back between two CUDA
    events on the launching stream.  Algorithmic bytes: 17.99 elements/pixel (what a statistics pass + a
    gradient pass must move; the single pass moves 13.33).  Reported for bf16 maps (BASELINE.json's loss-kernel metric) and
    for the variant TrainStep.g_step launches (fp32 maps as the heads emit them, + LS-G on the critic logits)."""
    import ctypes as C
    from octave_b200 import _lib, losses
    dev = torch.device("cuda")
    out = {}
    for tag, dt, with_lsg in (("loss_kernel", torch.bfloat16, False), ("loss_kernel_in_step", torch.float32, True)):
        es = 2 if dt == torch.bfloat16 else 4
        npx = B * H * W
        nsets = max(3, int(400e6 // (13.33 * es * npx)) + 1)
        sets = []
        for sd_ in range(nsets):
            g = torch.Generator(device=dev).manual_seed(sd_)
            agg = torch.randn(B, 2, H, W, device=dev, generator=g).to(dt)
            ys = torch.zeros(B, 2, H, W, device=dev, dtype=dt)
            ys[:, 0, ::37, :] = 1; ys[:, 1, 11::41, :] = 1
            att = [torch.softmax(torch.randn(B, 2, H >> k, W >> k, device=dev, generator=g), 1).to(dt) for k in range(5)]
            sets.append((agg, ys, att, torch.empty_like(agg), [torch.empty_like(a) for a in att]))
        fake = torch.randn(B, 1, device=dev) if with_lsg else None
        g_fake = torch.empty_like(fake) if with_lsg else None
        flags = _lib.LOSS_WPCE | _lib.LOSS_KLD | _lib.LOSS_FROM_LOGITS | (_lib.LOSS_LSG if with_lsg else 0)
        cfg = losses._LossCfg(flags, att_weights=[1.0] * 4, sum_weights=4.0)
        desc = losses._build_desc(cfg, sets[0][0], sets[0][2], None, fake)
        stats = torch.zeros(_lib.lib.octave_loss_fused_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)   # zero on entry, left zero by every evaluation
        outv = torch.empty(8, device=dev)
        lam = (C.c_float * 3)(1.0, 0.1, 0.1)
        sp = torch.cuda.current_stream().cuda_stream
        fns = []
        for agg, ys, att, g_y, g_a in sets:
            arr, garr = losses._ptr_array(att), losses._ptr_array(g_a)
            fns.append(lambda agg=agg, ys=ys, arr=arr, g_y=g_y, garr=garr: _lib.lib.octave_loss_fused(
                C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, losses._ptr(fake), lam, stats.data_ptr(), outv.data_ptr(), g_y.data_ptr(),
                garr, losses._ptr(g_fake), sp))
        for f in fns:
            assert f() == 0
        torch.cuda.synchronize()
        reps = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fns[i % nsets]()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = 17.99 * es * npx
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[tag] = {"kernel": f"loss_fused_kernel<{'bf16' if es == 2 else 'float'}, pyramid, co-resident> (ONE launch: label counts, KLD, WPCE, "
                              f"values and gradients{', LS-G' if with_lsg else ''}, finalisation)", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes": nbytes, "bytes_moved": 13.33 * es * npx, "us_per_evaluation": ms * 1e3,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this launch at B=32, 400^2, bf16
                    # (profiles/ncu_loss_fused_r02c.txt: 76.8 MB read + 13.6 MB written; most of the 68 MB of gradient writes are
                    # still in the 126 MB L2 when the kernel ends, so DRAM traffic is BELOW the 136.5 MB the kernel moves)
                    "traffic": 90.4e6 if (B, H, W, es, with_lsg) == (32, 400, 400, 2, False) else None,
                    "peak_source": peaks["source"], "maps": f"[{B},2,{H},{W}] {'bf16' if es == 2 else 'fp32'}, {nsets} rotating sets"}
        del sets, fns
        torch.cuda.empty_cache()
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
        # optional cap of the CTAs NCCL may use per collective (they share SMs with the persistent convolution kernels)
        if args.nccl_max_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))
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
    ts = TrainStep(net, distributed=world > 1, bucket_bytes=args.bucket_mb << 20, grad_dtype=None if args.grad_dtype == "auto" else args.grad_dtype)
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
                ts.predraw()                    # ... and so do the next step's CPU noise draws (before the blocking read below)
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
                   "optimizer": "SGD(momentum) as ONE multi-tensor launch per module (octave_optim_multi) + ONE launch re-packing the bf16 operands of all conv weights (SURVEY.md §8f.1)"},
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
                dev_batch = tuple(t.to(dev) if torch.is_tensor(t) else [r.to(dev) for r in t] for t in host[0])
                fams, eager_kernel_ms = family_rooflines(ts, dev_batch, peaks, ms_step)
                line["kernel_families"] = fams
                line["kernel_time_of_one_eager_step_ms"] = eager_kernel_ms
                dom = next(f for f in fams if f["bound"] in ("tensor", "hbm"))
                line["roofline"] = {"kernel": dom["family"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                                    "unit": dom["unit"], "frac": dom["frac"], "traffic": None,
                                    "share_of_kernel_time": dom["share_of_kernel_time"], "launches_per_step": dom["launches"],
                                    "how": "largest kernel family of the step by summed duration; CUDA events around each launch of one eager "
                                           "step on the launching stream; peak = " + ("sustained bf16" if dom["bound"] == "tensor" else "measured HBM copy")
                                           + " (" + peaks["source"] + ")"}
                line["best_kernel"] = best_conv_roofline(peaks, min(B, 32), H, W)
                line.update(loss_rooflines(peaks, min(B, 32), H, W))
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
        # tear-down: drop the captured graph (it holds the NCCL work of the all-reduces) before the communicator goes
        ts._graph = None
        ts._sout = None
        torch.cuda.synchronize()
        import gc
        gc.collect()
        try:
            dist.destroy_process_group()
        except Exception as e:          # pragma: no cover - never fatal for a finished benchmark
            sys.stderr.write(f"[bench] destroy_process_group: {e!r}\n")


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path — its modules cannot travel to the GPU box
    (/root/reference is absent there and the repo has no installer), so the oracle port runs, on all host cores.  Each step
    is one G+D iteration on a bounded sample (batch 2) of the arm's config; exactly `steps` steps after `warmup` are timed."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    B, H, W = CONFIGS[args.config]
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    if args.ref_budget_s > 0:
        # keep the arm within its time budget on slow hosts: ~1 s per batch-2 step at 400^2 on 16 cores
        est = 1.0 * (H * W) / 160000.0
        steps = max(1, min(steps, int(args.ref_budget_s / est) - warmup))
    cb = run_cpu(H, W, 2, steps, warmup)
    line = {"impl": "reference", "metric": f"train images/s, {H}x{W} synthetic OCTA (adversarial step: G-step + D-step)",
            "value": cb["value"], "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": steps,
            "warmup": warmup, "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{args.config}: OctaScribbleNet G+D training step, {H}x{W}; each step is a bounded sample of batch 2 "
                                   f"(BASELINE.json configs[0] is this step at 304x304, batch 2: --config c1)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: cap the number of timed steps so that the run fits this many seconds (0 = no cap)")
    ap.add_argument("--comparator", action="store_true", help="also time the reference arithmetic on stock torch CUDA ops (cuDNN) on this GPU")
    ap.add_argument("--bucket-mb", type=int, default=25, help="gradient all-reduce bucket size (data-parallel runs)")
    ap.add_argument("--grad-dtype", default="auto", choices=["auto", "fp32", "bf16"],
                    help="dtype of the all-reduced gradient buckets (auto: bf16 with --dtype bf16, fp32 otherwise)")
    ap.add_argument("--nccl-max-ctas", type=int, default=0, help="cap of NCCL CTAs per collective (0 = NCCL default; measured at N=2: "
                    "default 47.2 ms/step, cap 8 48.0, one unbucketed all-reduce 48.0 - profiles/ddp_ab_r02.log)")
    ap.add_argument("--no-graph", action="store_true", help="launch the ~1100 kernels of a step eagerly instead of replaying one CUDA graph")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
