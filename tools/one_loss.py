"""Launch the single-pass loss (labels pre-pass + fused kernel) a few times: target of ncu captures.  args: B H [bf16|fp32]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import _lib, losses
B, H = int(sys.argv[1]), int(sys.argv[2])
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "fp32") else torch.bfloat16
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
agg = torch.randn(B, 2, H, H, device=dev, generator=g).to(dt)
ys = torch.zeros(B, 2, H, H, device=dev, dtype=dt)
ys[:, 0, ::37, :] = 1; ys[:, 1, 11::41, :] = 1
att = [torch.softmax(torch.randn(B, 2, H >> k, H >> k, device=dev, generator=g), 1).to(dt) for k in range(5)]
g_y, g_a = torch.empty_like(agg), [torch.empty_like(a) for a in att]
cfg = losses._LossCfg(_lib.LOSS_WPCE | _lib.LOSS_KLD | _lib.LOSS_FROM_LOGITS, att_weights=[1.0] * 4, sum_weights=4.0)
desc = losses._build_desc(cfg, agg, att, None, None)
stats = torch.zeros(_lib.lib.octave_loss_fused_stats_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)   # zero on entry, left zero by every evaluation
outv = torch.empty(8, device=dev)
lam = (C.c_float * 3)(1.0, 0.1, 0.1)
arr, garr = losses._ptr_array(att), losses._ptr_array(g_a)
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    rc = _lib.lib.octave_loss_fused(C.byref(desc), agg.data_ptr(), ys.data_ptr(), arr, None, lam, stats.data_ptr(), outv.data_ptr(),
                                    g_y.data_ptr(), garr, None, sp)
    assert rc == 0
torch.cuda.synchronize()
print("ok", outv.tolist())
