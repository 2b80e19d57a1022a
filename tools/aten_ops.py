"""Which torch (ATen) ops still launch kernels inside one training step, by source line of this package.
usage: python tools/aten_ops.py [B H]"""
import collections, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, synth
from octave_b200.model import OctaScribbleNet
from octave_b200.train import TrainStep
from torch.utils._python_dispatch import TorchDispatchMode
import traceback

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 400
config.set_compute_dtype("bf16"); config.nan_check = False
torch.manual_seed(0)
net = OctaScribbleNet(torch.Size((B, 3, H, H)), torch.Size((B, 2, H, H)), True, False).cuda().train()
ts = TrainStep(net)
x, ys, _ = synth.octa_batch(B, H, H, seed=0, n_ridges=8)
real = [r.cuda() for r in synth.mask_pyramid(B, H, H, n_ridges=8)]
x, ys = x.cuda(), ys.cuda()
for _ in range(3):
    ts.step(x, ys, real)
torch.cuda.synchronize()

SKIP = ("aten.view", "aten.reshape", "aten._unsafe_view", "aten.detach", "aten.slice", "aten.select", "aten.as_strided", "aten.t.",
        "aten.transpose", "aten.permute", "aten.expand", "aten.unsqueeze", "aten.squeeze", "aten.alias", "aten.empty", "aten.set_",
        "aten.is_", "aten.size", "aten.stride", "aten._local_scalar_dense", "aten.record_stream", "aten.lift_fresh", "aten.unbind", "aten.split")
count = collections.Counter()

class Mode(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not name.startswith(SKIP):
            site = "?"
            for fr in reversed(traceback.extract_stack()):
                if "octave_b200" in fr.filename and "aten_ops" not in fr.filename:
                    site = f"{os.path.basename(fr.filename)}:{fr.lineno}"
                    break
            count[(name, site)] += 1
        return func(*args, **(kwargs or {}))

with Mode():
    ts.step(x, ys, real)
torch.cuda.synchronize()
tot = 0
for (name, site), n in sorted(count.items(), key=lambda kv: -kv[1]):
    print(f"{n:4d}  {name:40s} {site}")
    tot += n
print("total dispatched ops that may launch:", tot)
