"""Per-shape table of the conv entry points in one training step: time, FLOPs, TFLOP/s, bytes, GB/s."""
import os, re, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, profiler, synth
from octave_b200.model import OctaScribbleNet
from octave_b200.train import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 400
config.set_compute_dtype("bf16"); config.nan_check = False
config.overlap_wgrad = False        # one stream: per-launch event times are not stretched by concurrent kernels
torch.manual_seed(0)
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
ts = TrainStep(net)
x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=8)
real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=8)]
x, ys = x.cuda(), ys.cuda()
for _ in range(2):
    ts.step(x, ys, real)
profiler.enable()
ts.step(x, ys, real); profiler.reset()
ts.step(x, ys, real)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for k, e0, e1, _d in profiler._records:
    agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1)
rows = []
for k, (n, ms) in agg.items():
    m = re.match(r"(octave_conv_\w+)\[B(\d+) (\d+)x(\d+) (\d+)->(\d+) k(\d+) g(\d+) m(\d+)\]", k)
    if not m:
        continue
    name, b, h, w, ci, co, ks, g, mode = m.group(1), *map(int, m.groups()[1:])
    pix = b * h * w
    co_eff = co * (4 if mode == 1 else 1)
    flops = 2.0 * pix * co_eff * (ci // g) * ks * ks
    byts = 2.0 * pix * (ci + co_eff)
    t = ms / n
    rows.append((ms, n, name.replace("octave_conv_", ""), f"{h}x{w} {ci}->{co} k{ks} g{g} m{mode}", t, flops / t / 1e9, byts / t / 1e6))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"conv total {tot:.1f} ms")
for ms, n, name, shp, t, tf, gb in rows:
    print(f"{ms:7.3f} ms n={n:2d} {name:9s} {shp:34s} {t*1e3:8.1f} us  {tf:7.1f} TF/s  {gb:7.0f} GB/s(min traffic)")
