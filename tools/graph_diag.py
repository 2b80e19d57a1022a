import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_train_graph_gpu import _make, _data
x, ys, real = _data()
def run(mode, n):
    net, ts = _make(0)
    w0 = {k: v.clone() for k, v in net.state_dict().items()}
    torch.manual_seed(99)
    if mode == "eager":
        for _ in range(n): ts.step(x, ys, real)
    else:
        for _ in range(n - 2): ts.step_graphed(x, ys, real)
        assert ts.graph_error is None, ts.graph_error
    return w0, {k: v.clone() for k, v in net.state_dict().items()}
w0, a = run("eager", 4)
_, a2 = run("eager", 4)
_, b = run("graph", 4)
keys = [k for k in a if a[k].is_floating_point() and "running" not in k and not k.startswith("segmentor.linear_head_") and not k.endswith(("_u", "_v"))]
def rel(p, q):
    num = sum(float(((p[k] - q[k]).float() ** 2).sum()) for k in keys) ** 0.5
    den = sum(float(((p[k] - w0[k]).float() ** 2).sum()) for k in keys) ** 0.5
    return num / den
print("eager vs eager  |dW diff| / |dW|:", rel(a, a2))
print("eager vs graph  |dW diff| / |dW|:", rel(a, b))
worst = sorted(((float((a[k] - b[k]).norm() / (a[k] - w0[k]).norm().clamp_min(1e-12)), float((a[k] - a2[k]).norm() / (a[k] - w0[k]).norm().clamp_min(1e-12)), k) for k in keys), reverse=True)[:8]
for w in worst: print("graph %.3f eager-eager %.3f %s" % w)
