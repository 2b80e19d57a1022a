import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import ops
from octave_b200.ops import Act
B, H, C = 32, int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
x = Act(torch.randn(B, H, H, C, device=dev).bfloat16(), B, H, H, C)
dy = Act(torch.randn(B, H, H, C, device=dev).bfloat16(), B, H, H, C)
mi = torch.zeros(2 * C, device=dev); mi[C:] = 1.0
ab = torch.ones(2 * C, device=dev); gamma = torch.ones(C, device=dev)
for _ in range(3):
    ops.chan_stats(x)
    ops.bn_bwd(dy, None, x, mi, gamma, True, out=dy, relu_ab=ab)
torch.cuda.synchronize()
print("ok")
