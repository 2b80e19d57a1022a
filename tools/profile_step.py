"""Per-entry-point GPU time of one training step (CUDA events; warm)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, profiler, synth
from octave_b200.model import OctaScribbleNet
from octave_b200.train import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 400
config.set_compute_dtype("bf16"); config.nan_check = False
torch.manual_seed(0)
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
ts = TrainStep(net)
x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=8)
real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=8)]
x, ys = x.cuda(), ys.cuda()
for _ in range(2):
    ts.step(x, ys, real)
torch.cuda.synchronize()
t0 = time.perf_counter(); ts.step(x, ys, real); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"wall per step (no profiler): {1e3 * (t1 - t0):.1f} ms; max mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
profiler.enable()
ts.step(x, ys, real); profiler.reset()
t0 = time.perf_counter(); ts.step(x, ys, real); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"wall per step (profiler on): {1e3 * (t1 - t0):.1f} ms")
print(profiler.report(30))
print(profiler.report(45, by_shape=True))
# host-only cost: time the python side with the GPU idle at the end
