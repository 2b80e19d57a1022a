"""A/B of the forward kernels with and without the slab epilogue (swizzled staging, one more pipeline stage):
  256-column tiles: OCTAVE_FWD_STAGES=3 vs 4 (default 4)        suite "wide"
  128-column tiles: OCTAVE_FWD128_STAGES=5 vs 6 (default 5)     suite "mid"
usage: slab_probe.py save|compare <file> [wide|mid]   (one process per setting: the env switches are read once)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import ops
from octave_b200.ops import Act, ConvSpec
mode, path = sys.argv[1], sys.argv[2]
dev = torch.device("cuda")
suite = sys.argv[3] if len(sys.argv) > 3 else "wide"
if suite == "wide":   # B, H, W, cin, cout, k, fused statistics, bias + ReLU
    CASES = [(32, 100, 100, 512, 256, 3, True, False), (60, 37, 41, 256, 512, 3, True, False), (32, 50, 50, 1024, 512, 3, False, True)]
else:
    CASES = [(32, 100, 100, 64, 256, 1, True, False), (32, 50, 50, 512, 128, 3, True, False), (8, 37, 41, 128, 128, 3, True, False),
             (32, 100, 100, 256, 128, 3, False, True)]
outs, times = [], []
for (B, H, W, cin, cout, k, stats, bias_relu) in CASES:
    g = torch.Generator(device=dev).manual_seed(B + cin)
    x = Act(torch.randn(B, H, W, cin, device=dev, generator=g).bfloat16(), B, H, W, cin)
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device=dev, generator=g) * 0.02)
    b = torch.nn.Parameter(torch.randn(cout, device=dev, generator=g)) if bias_relu else None
    spec = ConvSpec(w, b, cin, cout, k, 1, k // 2, 1)
    run = (lambda: ops.conv_fwd(x, spec, want_stats=True)) if stats else (lambda: (ops.conv_fwd(x, spec, act=1), None))
    for _ in range(2): y, st = run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): y, st = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    times.append(ms)
    outs.append((y.buf.cpu(), None if st is None else st.cpu()))
    print(f"stages256={os.environ.get('OCTAVE_FWD_STAGES', '4')} stages128={os.environ.get('OCTAVE_FWD128_STAGES', '5')} B{B} {H}x{W} {cin}->{cout} k{k}: {ms*1e3:.1f} us  {2.0*B*H*W*cin*cout*k*k/ms/1e9:.0f} TF/s")
if mode == "save":
    torch.save(outs, path)
else:
    ref = torch.load(path)
    for i, ((y, st), (yr, sr)) in enumerate(zip(outs, ref)):
        ok_y = torch.equal(y, yr)
        ok_s = True if st is None else torch.equal(st, sr)
        md = float((y.float() - yr.float()).abs().max())
        print(f"case {i}: y bit-exact {ok_y} (max diff {md}), stats bit-exact {ok_s}" + ("" if st is None or ok_s else f" max rel {float(((st - sr).abs() / sr.abs().clamp_min(1e-9)).max())}"))
