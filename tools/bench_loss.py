"""Fused loss kernel (K9) alone: forward statistics launch and gradient launch timed separately through the C-ABI,
L2 flushed between iterations; GB/s = algorithmic bytes (fwd 6.664, bwd 11.328 elements/pixel) / time."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import _lib, losses
dev = torch.device("cuda")
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
for (B, H, dt) in [(32, 400, torch.bfloat16), (64, 304, torch.bfloat16), (8, 1024, torch.bfloat16), (32, 400, torch.float32)]:
    g = torch.Generator(device=dev).manual_seed(0)
    agg = torch.randn(B, 2, H, H, device=dev, generator=g).to(dt)
    ys = torch.zeros(B, 2, H, H, device=dev, dtype=dt)
    # scribble-like labels: sparse horizontal strokes (~3% per class)
    ys[:, 0, ::37, :] = 1; ys[:, 1, 11::41, :] = 1
    att = [torch.softmax(torch.randn(B, 2, H >> k, H >> k, device=dev, generator=g), 1).to(dt) for k in range(5)]
    cfg = losses._LossCfg(_lib.LOSS_WPCE | _lib.LOSS_KLD | _lib.LOSS_FROM_LOGITS, att_weights=[1.0] * 4, sum_weights=4.0)
    desc = losses._build_desc(cfg, agg, att, None, None)
    stats = torch.empty(_lib.lib.octave_loss_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
    outv = torch.empty(8, device=dev); gs = torch.ones(8, device=dev)
    g_y = torch.empty_like(agg); g_a = [torch.empty_like(a) for a in att]
    arr, garr = losses._ptr_array(att), losses._ptr_array(g_a)
    sp = torch.cuda.current_stream().cuda_stream
    fwd = lambda: _lib.lib.octave_loss_fwd(C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, None, stats.data_ptr(), outv.data_ptr(), sp)
    bwd = lambda: _lib.lib.octave_loss_bwd(C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, None, stats.data_ptr(), gs.data_ptr(),
                                           g_y.data_ptr(), garr, None, None, sp)
    def t(fn, reps=10):
        for _ in range(3): fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    es = agg.element_size(); npx = B * H * H
    tf, tb = t(fwd), t(bwd)
    print(f"B{B} {H}x{H} {dt}: fwd {tf*1e3:6.1f} us {6.664*es*npx/tf/1e6:6.0f} GB/s | bwd {tb*1e3:6.1f} us {11.328*es*npx/tb/1e6:6.0f} GB/s | "
          f"fwd+bwd {17.99*es*npx/(tf+tb)/1e6:6.0f} GB/s")
