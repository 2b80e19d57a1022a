"""Probe / A-B check of the narrow-layer halo conv kernel against the generic tcgen05 kernel (same inputs, same packs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import ops
from octave_b200.ops import Act, ConvSpec, lib
dev = torch.device("cuda")
torch.manual_seed(0)

def run(B, H, W, cin, cout, groups, mode, want_stats=True):
    x = Act(torch.randn(B, H, W, cin, device=dev).bfloat16(), B, H, W, cin)
    w = torch.nn.Parameter(torch.randn(cout, cin // groups, 3, 3, device=dev) * 0.05)
    spec = ConvSpec(w, None, cin, cout, 3, 1, 1, groups)
    y, st = ops.conv_fwd(x, spec, want_stats=want_stats)
    dx = ops.conv_dgrad(y, spec, H, W)
    dx2 = ops.conv_dgrad(y, spec, H, W, out=Act(dx.buf.clone(), B, H, W, cin), accumulate=True)
    dw, _ = ops.conv_wgrad(x, y, spec)
    torch.cuda.synchronize()
    return y.buf.float(), (st.clone() if st is not None else None), dx.buf.float(), dx2.buf.float(), dw.float()

def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

shapes = [(2, 64, 64, 64, 32, 1), (2, 40, 52, 32, 64, 1), (3, 33, 47, 64, 128, 2), (2, 100, 100, 128, 64, 2), (1, 16, 16, 32, 32, 1), (2, 50, 50, 64, 64, 1)]
MODES = [int(a) for a in sys.argv[1:]] or [0]
for mode in MODES:
    print(f"--- base_off_mode {mode}")
    for shp in shapes:
        torch.manual_seed(1)
        lib.octave_conv_halo_config(0, 0)
        ref = run(*shp, mode)
        torch.manual_seed(1)
        lib.octave_conv_halo_config(1, mode)
        try:
            got = run(*shp, mode)
        except Exception as e:
            print(shp, "ERROR", e)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("  sync:", e2)
            sys.exit(1)
        errs = []
        for a, b in zip(ref, got):
            if a is None: continue
            a = a.double(); b = b.double()
            errs.append(float((a - b).abs().max() / (a.abs().max() + 1e-9)))
        print(shp, "rel max err (y, stats, dx, dx_acc, dw):", ["%.2e" % e for e in errs])
HM = MODES[0]
for (B, H, cin, cout, g) in [(32, 400, 64, 32, 1), (32, 400, 32, 64, 1), (32, 200, 32, 64, 1), (32, 200, 64, 128, 2), (32, 200, 32, 32, 1), (32, 100, 64, 64, 1)]:
    x = Act(torch.randn(B, H, H, cin, device=dev).bfloat16(), B, H, H, cin)
    w = torch.nn.Parameter(torch.randn(cout, cin // g, 3, 3, device=dev) * 0.05)
    spec = ConvSpec(w, None, cin, cout, 3, 1, 1, g)
    y = Act.empty(B, H, H, cout, torch.bfloat16, dev)
    res = []
    for en in (0, 1):
        lib.octave_conv_halo_config(en, HM)
        ms = timeit(lambda: ops.conv_fwd(x, spec, out=y, want_stats=True))
        res.append(ms)
    byts = 2.0 * B * H * H * (cin + cout)
    resw = []
    for en in (0, 1):
        lib.octave_conv_halo_config(en, HM)
        resw.append(timeit(lambda: ops.conv_wgrad(x, y, spec)))
    print(f"B{B} {H}x{H} {cin}->{cout} g{g}: fwd generic {res[0]*1e3:7.1f} us  halo {res[1]*1e3:7.1f} us  ({byts/res[1]/1e6:.0f} GB/s) | "
          f"wgrad generic {resw[0]*1e3:7.1f} us  halo {resw[1]*1e3:7.1f} us")
