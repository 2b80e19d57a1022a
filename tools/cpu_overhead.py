"""Host-side cost of one step: run a tiny problem (GPU work negligible) and time wall clock."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, synth, _lib
from octave_b200.model import OctaScribbleNet
from octave_b200.train import TrainStep
config.set_compute_dtype("bf16"); config.nan_check = False
for B, H in ((2, 64), (32, 400)):
    torch.manual_seed(0)
    net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
    ts = TrainStep(net)
    x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=4)
    real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=4)]
    x, ys = x.cuda(), ys.cuda()
    for _ in range(3):
        ts.step(x, ys, real)
    torch.cuda.synchronize()
    n0 = _lib.lib.octave_launch_count()
    t0 = time.perf_counter()
    for _ in range(3):
        ts.step(x, ys, real)
    t_enq = time.perf_counter()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"B={B} H={H}: wall {1e3*(t1-t0)/3:.1f} ms/step, enqueue-only {1e3*(t_enq-t0)/3:.1f} ms/step, launches/step {(_lib.lib.octave_launch_count()-n0)//3}")
    del net, ts
import cProfile, pstats
B, H = 2, 64
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
ts = TrainStep(net)
x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=4)
real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=4)]
x, ys = x.cuda(), ys.cuda()
for _ in range(2): ts.step(x, ys, real)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): ts.step(x, ys, real)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
