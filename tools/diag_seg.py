import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_segmentor_gpu import _build, _oracle_step, _cuda_step, l2err
from tests import synth
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
size = int(sys.argv[3]) if len(sys.argv) > 3 else 112
net, sd = _build(mode)
x, ys, _ = synth.octa_batch(B, size, size, seed=size)
att_o, agg_o, x4_o, wp_o, kl_o, g_o, st = _oracle_step(sd, x, ys)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
att_d, agg_d, x4_d, wp_d, kl_d, g_d, _ = _oracle_step(sd64, x.double(), ys.double())
att, agg, x4, wp, kl = _cuda_step(net, x, ys)
print("agg", l2err(agg, agg_d), "oracle32:", l2err(agg_o, agg_d), " x4", l2err(x4, x4_d), l2err(x4_o, x4_d))
for a, b, c in zip(att, att_d, att_o):
    print(" att", tuple(a.shape), l2err(a, b), "oracle32:", l2err(c, b))
params = dict(net.named_parameters())
rows = []
for k, gd in g_d.items():
    if gd is None: continue
    rows.append((l2err(params[k].grad, gd), l2err(g_o[k], gd), float(gd.abs().max()), k))
gs = max(r[2] for r in rows)
rows = [r for r in rows if r[2] > 1e-6 * gs]
order = [k for k in g_d if g_d[k] is not None]
byname = {r[3]: r for r in rows}
for k in order:
    if k in byname and (k.endswith("weight") and ("conv" in k or "up." in k or "fc." in k or "downsample.0" in k or "downsample.1.w" in k)):
        r = byname[k]
        print("%.3e  oracle32 %.3e  ratio %6.1f  %s" % (r[0], r[1], r[0] / max(r[1], 1e-12), k))
import statistics
print("median mine", statistics.median(r[0] for r in rows), "median oracle32", statistics.median(r[1] for r in rows))
