"""Run-to-run noise of eval-mode bf16 gradients (calibrates the tolerance of tests/test_ddp_nccl_gpu.py) on ONE GPU."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_ddp_nccl_gpu import _build, _shard_grads
from octave_b200 import synth
net = _build()
sh = synth.octa_batch(2, 96, 96, seed=40, n_ridges=8)[:2]
outs = []
for _ in range(3):
    _shard_grads(net, *sh)
    outs.append({n: p.grad.detach().float().clone() for n, p in net.named_parameters() if p.grad is not None})
for i in (1, 2):
    errs = sorted(((float((outs[i][n] - outs[0][n]).norm() / outs[0][n].norm().clamp_min(1e-20)), n) for n in outs[0]), reverse=True)
    print("run", i, "vs 0: worst", errs[:4], "median", errs[len(errs) // 2][0])
