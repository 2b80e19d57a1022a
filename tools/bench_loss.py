"""Fused loss kernel (K9) alone, through the C-ABI.  Timing: NSETS rotating input sets whose total footprint exceeds the
126 MB L2, REPS launches queued back to back between two CUDA events on the launching stream (no host gaps, no L2 hits
from the previous launch).  GB/s = algorithmic bytes of SURVEY.md 8d (17.99 elements/pixel fwd+bwd) / time."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import _lib, losses
dev = torch.device("cuda")


def make_set(B, H, dt, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    agg = torch.randn(B, 2, H, H, device=dev, generator=g).to(dt)
    ys = torch.zeros(B, 2, H, H, device=dev, dtype=dt)
    ys[:, 0, ::37, :] = 1; ys[:, 1, 11::41, :] = 1          # scribble-like labels: sparse strokes (~5 % of the pixels)
    att = [torch.softmax(torch.randn(B, 2, H >> k, H >> k, device=dev, generator=g), 1).to(dt) for k in range(5)]
    return agg, ys, att, torch.empty_like(agg), [torch.empty_like(a) for a in att]


def time_queue(fns, reps):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (B, H, dt) in [(32, 400, torch.bfloat16), (64, 304, torch.bfloat16), (8, 1024, torch.bfloat16), (32, 400, torch.float32)]:
    es = 2 if dt == torch.bfloat16 else 4
    npx = B * H * H
    nsets = max(3, int(400e6 // (13.33 * es * npx)) + 1)
    sets = [make_set(B, H, dt, s) for s in range(nsets)]
    cfg = losses._LossCfg(_lib.LOSS_WPCE | _lib.LOSS_KLD | _lib.LOSS_FROM_LOGITS, att_weights=[1.0] * 4, sum_weights=4.0)
    desc = losses._build_desc(cfg, sets[0][0], sets[0][2], None, None)
    stats = torch.zeros(_lib.lib.octave_loss_fused_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)   # zero on entry, left zero by every evaluation
    stats2 = torch.empty(_lib.lib.octave_loss_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)       # two-pass form: its own workspace
    outv = torch.empty(8, device=dev); gs = torch.ones(8, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    lam = (C.c_float * 3)(1.0, 0.1, 0.1)
    fw, bw, fu = [], [], []
    for agg, ys, att, g_y, g_a in sets:
        arr, garr = losses._ptr_array(att), losses._ptr_array(g_a)
        fw.append(lambda agg=agg, ys=ys, arr=arr: _lib.lib.octave_loss_fwd(C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, None,
                                                                           stats2.data_ptr(), outv.data_ptr(), sp))
        bw.append(lambda agg=agg, ys=ys, arr=arr, g_y=g_y, garr=garr: _lib.lib.octave_loss_bwd(
            C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, None, stats2.data_ptr(), gs.data_ptr(), g_y.data_ptr(), garr, None, None, sp))
        fu.append(lambda agg=agg, ys=ys, arr=arr, g_y=g_y, garr=garr: _lib.lib.octave_loss_fused(
            C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, lam, stats.data_ptr(), outv.data_ptr(), g_y.data_ptr(), garr, None, sp))
    assert fu[0]() == 0
    tf, tb, tfu = time_queue(fw, 40), time_queue(bw, 40), time_queue(fu, 40)
    print(f"B{B} {H}x{H} {dt} ({nsets} rotating sets): single pass {tfu*1e3:6.1f} us = {17.99*es*npx/tfu/1e6:6.0f} GB/s algorithmic "
          f"({13.33*es*npx/tfu/1e6:6.0f} GB/s moved) | two-pass fwd {tf*1e3:6.1f} us + bwd {tb*1e3:6.1f} us = {17.99*es*npx/(tf+tb)/1e6:6.0f} GB/s")
