"""Launch one tensor-core conv shape a few times (target of ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import ops
from octave_b200.ops import Act, ConvSpec
B, H, cin, cout, k = [int(a) for a in sys.argv[1:6]]
groups = int(sys.argv[6]) if len(sys.argv) > 6 else 1
mode = sys.argv[7] if len(sys.argv) > 7 else "fwd"
dev = torch.device("cuda")
x = Act(torch.randn(B, H, H, cin, device=dev).bfloat16(), B, H, H, cin)
w = torch.nn.Parameter(torch.randn(cout, cin // groups, k, k, device=dev) * 0.05)
spec = ConvSpec(w, None, cin, cout, k, 1, k // 2, groups)
y = Act.empty(B, H, H, cout, torch.bfloat16, dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run():
    if mode == "fwd":
        ops.conv_fwd(x, spec, out=y, want_stats=True)
    elif mode == "wgrad":
        ops.conv_wgrad(x, y, spec)
for _ in range(3): run()
torch.cuda.synchronize(); e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = 2.0 * B * H * H * cout * (cin // groups) * k * k
print(f"{mode} B{B} {H}x{H} {cin}->{cout} k{k} g{groups}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TF/s  {2.0*B*H*H*(cin+cout)/ms/1e6:.0f} GB/s(min)")
