"""Can the segmentor forward+backward be captured in a CUDA graph, and what does replay save over eager launches?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, synth
from octave_b200.model import OctaScribbleNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 400
config.set_compute_dtype("bf16"); config.nan_check = False
torch.manual_seed(0)
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
seg = net.segmentor
x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=8)
x = x.cuda()
params = [p for n, p in seg.named_parameters() if not n.startswith("linear_head_")]

def fwd_bwd():
    att, agg, x4 = seg(x)
    loss = agg.float().mean() + sum(a.float().mean() for a in att)
    gs = torch.autograd.grad(loss, params, allow_unused=True)
    return loss, gs

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print(f"eager fwd+bwd: {timeit(fwd_bwd):.2f} ms")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): fwd_bwd()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        loss, gs = fwd_bwd()
    torch.cuda.synchronize()
    print(f"graph replay fwd+bwd: {timeit(g.replay):.2f} ms; loss {float(loss):.5f}")
except Exception as e:
    print("capture failed:", repr(e)[:600])
