"""CUDA-event time of the phases of one adversarial step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, synth
from octave_b200.model import OctaScribbleNet
from octave_b200.train import TrainStep
B, H = 32, 400
config.set_compute_dtype("bf16"); config.nan_check = False
torch.manual_seed(0)
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
ts = TrainStep(net)
x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=8)
real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=8)]
x, ys = x.cuda(), ys.cuda()
for _ in range(3): ts.step(x, ys, real)
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {}
def mark(name, e0, e1): acc.setdefault(name, []).append((e0, e1))
for _ in range(3):
    e = [ev() for _ in range(10)]
    ts.opt_g.zero_grad(set_to_none=True); ts._set_d_grad(False)
    e[0].record(); att, agg, _ = net.segmentor(x)
    e[1].record(); y_fake = net.discriminator(att)
    e[2].record(); res = ts.loss(agg, ys, att, y_fake); total = res['supervised'] + 0.1 * res['divergence'] + 0.1 * res['generator']
    e[3].record(); total.backward()
    e[4].record(); ts.opt_g.step(); ts._set_d_grad(True)
    e[5].record(); ts.opt_d.zero_grad(set_to_none=True)
    dl = ts.lsd(net.discriminator(list(real)), net.discriminator([f.detach() for f in att]))
    e[6].record(); dl.backward()
    e[7].record(); ts.opt_d.step()
    e[8].record()
    torch.cuda.synchronize()
    for i, n in enumerate(["seg fwd", "D fwd (fake, G-step)", "fused loss fwd", "G backward (loss+D+seg)", "opt_g.step", "D fwd x2 + LSD", "D backward", "opt_d.step"]):
        acc.setdefault(n, []).append(e[i].elapsed_time(e[i + 1]))
for n, v in acc.items():
    print(f"{n:28s} {sorted(v)[len(v)//2]:8.2f} ms")
