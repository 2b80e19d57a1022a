"""Inference throughput of ResnestUNet.predict (eval, no_grad, bf16): conv+BN folded vs separate BN pass."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, network
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 400
config.set_compute_dtype("bf16")
torch.manual_seed(0)
net = network.ResnestUNet(2, False).cuda().eval()
x = torch.randn(B, 3, S, S, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fold in (True, False, True):
    config.fold_bn_inference = fold
    with torch.no_grad():
        for _ in range(3): net.predict(x, method='one-hot')
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): net.predict(x, method='one-hot')
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"predict B{B} {S}x{S} bf16 fold_bn={fold}: {ms:.2f} ms  {B/ms*1e3:.0f} img/s  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
