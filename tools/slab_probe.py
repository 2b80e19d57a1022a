"""A/B of the 256-column forward kernel: OCTAVE_FWD_STAGES=3 (padded staging) vs 4 (slab epilogue, swizzled staging).
usage: slab_probe.py save|compare <file>   (one process per setting: the env switch is read once)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import ops
from octave_b200.ops import Act, ConvSpec
mode, path = sys.argv[1], sys.argv[2]
dev = torch.device("cuda")
CASES = [(32, 100, 100, 512, 256, True, False), (60, 37, 41, 256, 512, True, False), (32, 50, 50, 1024, 512, False, True)]
outs, times = [], []
for (B, H, W, cin, cout, stats, bias_relu) in CASES:
    g = torch.Generator(device=dev).manual_seed(B + cin)
    x = Act(torch.randn(B, H, W, cin, device=dev, generator=g).bfloat16(), B, H, W, cin)
    w = torch.nn.Parameter(torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.02)
    b = torch.nn.Parameter(torch.randn(cout, device=dev, generator=g)) if bias_relu else None
    spec = ConvSpec(w, b, cin, cout, 3, 1, 1, 1)
    run = (lambda: ops.conv_fwd(x, spec, want_stats=True)) if stats else (lambda: (ops.conv_fwd(x, spec, act=1), None))
    for _ in range(2): y, st = run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): y, st = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    times.append(ms)
    outs.append((y.buf.cpu(), None if st is None else st.cpu()))
    print(f"stages={os.environ.get('OCTAVE_FWD_STAGES', '3')} B{B} {H}x{W} {cin}->{cout}: {ms*1e3:.1f} us  {2.0*B*H*W*cin*cout*9/ms/1e9:.0f} TF/s")
if mode == "save":
    torch.save(outs, path)
else:
    ref = torch.load(path)
    for i, ((y, st), (yr, sr)) in enumerate(zip(outs, ref)):
        ok_y = torch.equal(y, yr)
        ok_s = True if st is None else torch.equal(st, sr)
        md = float((y.float() - yr.float()).abs().max())
        print(f"case {i}: y bit-exact {ok_y} (max diff {md}), stats bit-exact {ok_s}" + ("" if st is None or ok_s else f" max rel {float(((st - sr).abs() / sr.abs().clamp_min(1e-9)).max())}"))
