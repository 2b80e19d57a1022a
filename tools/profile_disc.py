"""Per-entry-point GPU time of the critic alone: forward on 'fake' maps with gradient to the maps (G-step use) and a D-step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, profiler, synth, losses
from octave_b200.model import OctaScribbleNet
B, H = 32, 400
config.set_compute_dtype("bf16"); config.nan_check = False
torch.manual_seed(0)
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
D = net.discriminator
real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=8)]
fake = [torch.softmax(torch.randn(B, 2, H >> k, H >> k, device="cuda"), 1).requires_grad_() for k in range(5)]
lsd, lsg = losses.LSDiscriminatorialLoss(), losses.LSGeneratorLoss()
def g_part():
    for p in D.parameters(): p.requires_grad_(False)
    l = lsg(D(fake)); l.backward()
    for p in D.parameters(): p.requires_grad_(True)
def d_part():
    l = lsd(D(real), D([f.detach() for f in fake])); l.backward()
for _ in range(2): g_part(); d_part()
profiler.enable()
g_part(); d_part(); profiler.reset()
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record(); g_part(); e1.record(); d_part(); e2.record(); torch.cuda.synchronize()
print(f"G-part (D fwd + bwd to maps) {e0.elapsed_time(e1):.2f} ms; D-part (2 fwd + bwd to params) {e1.elapsed_time(e2):.2f} ms")
print(profiler.report(24))
print(profiler.report(30, by_shape=True))
