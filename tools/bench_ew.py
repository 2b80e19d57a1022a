"""Microbenchmark of the bandwidth-bound kernels: achieved GB/s per shape (algorithmic bytes / CUDA-event time)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import ops
from octave_b200.ops import Act
dev = torch.device("cuda")
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps

B = 32
for (H, C) in [(400, 32), (400, 64), (200, 64), (200, 128), (100, 256), (100, 64), (50, 512), (50, 128), (25, 1024), (25, 256), (13, 2048)]:
    x = Act(torch.randn(B, H, H, C, device=dev).bfloat16(), B, H, H, C)
    dy = Act(torch.randn(B, H, H, C, device=dev).bfloat16(), B, H, H, C)
    y = x.like()
    nbytes = B * H * H * C * 2
    mi = torch.zeros(2 * C, device=dev); mi[C:] = 1.0
    ab = torch.ones(2 * C, device=dev)
    gamma = torch.ones(C, device=dev)
    t_stats = timeit(lambda: ops.chan_stats(x))
    t_aff = timeit(lambda: ops.affine_act(x, ab, None, True, y))
    t_gap = timeit(lambda: ops.affine_act(x, ab, None, True, y, True))
    t_bwd = timeit(lambda: ops.bn_bwd(dy, y, x, mi, gamma, True, out=dy))
    t_bwd2 = timeit(lambda: ops.bn_bwd(dy, None, x, mi, gamma, True, out=dy, relu_ab=ab))
    t_relu = timeit(lambda: ops.relu_bwd(dy, y, y))
    print(f"{H:4d}x{H:<4d} C={C:5d} {nbytes/2**20:7.1f} MiB | stats {t_stats*1e3:7.1f} us {nbytes/t_stats/1e6:6.0f} GB/s | affine {t_aff*1e3:7.1f} us {2*nbytes/t_aff/1e6:6.0f} GB/s | "
          f"affine+gap {t_gap*1e3:7.1f} us {2*nbytes/t_gap/1e6:6.0f} | bn_bwd(reduce+apply) {t_bwd*1e3:7.1f} us {7*nbytes/t_bwd/1e6:6.0f} GB/s | selfmask {t_bwd2*1e3:7.1f} us {5*nbytes/t_bwd2/1e6:6.0f} | relu_bwd {t_relu*1e3:6.1f} us {3*nbytes/t_relu/1e6:6.0f}")
