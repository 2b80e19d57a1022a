"""GPU probe of the tcgen05 conv kernels with diagnostics (run under gpurun; writes gpurun_out/probe_tc.log)."""
import ctypes as C
import os
import sys
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import _lib  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/probe_tc_%s.log" % (sys.argv[1] if len(sys.argv) > 1 else "all"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    LOG.write(s + "\n")
    LOG.flush()


def stream():
    return torch.cuda.current_stream().cuda_stream


def desc(B, H, W, cin, cout, groups, k, x_ld=None, y_ld=None, mode=0, Hout=None, Wout=None, relu=0, out_dtype=1):
    d = _lib.ConvDesc()
    d.B, d.H, d.W, d.cin, d.cout, d.groups, d.ksize = B, H, W, cin, cout, groups, k
    d.stride, d.pad = 1, k // 2
    d.x_ld, d.x_coff = x_ld or cin, 0
    d.y_ld, d.y_coff = y_ld or cout, 0
    d.Hout, d.Wout = Hout or H, Wout or W
    d.mode, d.relu, d.in_dtype, d.out_dtype = mode, relu, 1, out_dtype
    return d


def report(name, out, ref):
    out = out.float().cpu(); ref = ref.float().cpu()
    err = (out - ref).abs()
    scale = ref.abs().max().item()
    bad = (err > 2e-2 * scale).float().mean().item()
    log(f"[{name}] max_err={err.max().item():.4e} ref_max={scale:.4e} rel={err.max().item()/max(scale,1e-9):.3e} frac_bad={bad:.4f}",
        "OK" if bad == 0 else "FAIL")
    return bad == 0


def conv_case(name, B, H, W, cin, cout, groups, k, relu=0, bias=True, identity=False):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, cin, H, W, generator=g).bfloat16()
    w = (torch.randn(cout, cin // groups, k, k, generator=g) / (cin // groups * k * k) ** 0.5).bfloat16()
    if identity:
        w.zero_()
        for i in range(min(cout, cin)):
            w[i, i, k // 2, k // 2] = 1.0
    b = torch.randn(cout, generator=g) if bias else None
    ref = F.conv2d(x.float(), w.float(), b, 1, k // 2, 1, groups)
    if relu:
        ref = F.relu(ref)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin // groups).contiguous().cuda()
    y = torch.full((B, H, W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    d = desc(B, H, W, cin, cout, groups, k, relu=relu)
    if not _lib.lib.octave_conv_tc_supported(C.byref(d)):
        log(f"[{name}] unsupported"); return False
    bc = b.cuda() if b is not None else None
    st = torch.full((2 * cout,), float("nan"), dtype=torch.float64, device="cuda")
    rc = _lib.lib.octave_conv_tc_fwd(C.byref(d), xn.data_ptr(), wp.data_ptr(), bc.data_ptr() if bc is not None else None,
                                     y.data_ptr(), st.data_ptr(), stream())
    torch.cuda.synchronize()
    if rc != 0:
        log(f"[{name}] rc={rc}"); return False
    ok = report(name, y.permute(0, 3, 1, 2), ref)
    yf = y.float().double()
    s_ref = torch.cat([yf.sum(dim=(0, 1, 2)), (yf * yf).sum(dim=(0, 1, 2))]).cpu()
    s_err = ((st.cpu() - s_ref).abs() / (s_ref.abs() + 1e-3 * s_ref.abs().max())).max().item()
    log(f"   fused stats rel err {s_err:.3e}", "OK" if s_err < 1e-4 else "FAIL")
    ok = ok and s_err < 1e-4
    if not ok:
        yo = y.float().cpu(); rf = ref.permute(0, 2, 3, 1)
        log("   out[0,0,0,:8] ", yo[0, 0, 0, :8].tolist()); log("   ref[0,0,0,:8] ", rf[0, 0, 0, :8].tolist())
        log("   out[0,0,1,:8] ", yo[0, 0, 1, :8].tolist()); log("   ref[0,0,1,:8] ", rf[0, 0, 1, :8].tolist())
        log("   nan frac", torch.isnan(yo).float().mean().item())
        e = (yo - rf).abs().amax(dim=(0, 3)); log("   err per (h,w) [first 4 rows]:", e[:4, :12].tolist())
        e = (yo - rf).abs().amax(dim=(0, 1, 2)); log("   err per channel[:32]:", e[:32].tolist())
    return ok


def dgrad_case(name, B, H, W, cin, cout, groups, k):
    g = torch.Generator().manual_seed(1)
    dy = torch.randn(B, cout, H, W, generator=g).bfloat16()
    w = (torch.randn(cout, cin // groups, k, k, generator=g) / (cout // groups * k * k) ** 0.5).bfloat16()
    x = torch.zeros(B, cin, H, W, requires_grad=True)
    F.conv2d(x, w.float(), None, 1, k // 2, 1, groups).backward(dy.float())
    ref = x.grad
    G, cg, og = groups, cin // groups, cout // groups
    wp = w.view(G, og, cg, k, k).flip(3, 4).permute(3, 4, 0, 2, 1).reshape(k * k, G * cg, og).contiguous().cuda()
    dyn = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dx = torch.full((B, H, W, cin), float("nan"), dtype=torch.bfloat16, device="cuda")
    d = desc(B, H, W, cout, cin, groups, k)
    if not _lib.lib.octave_conv_tc_supported(C.byref(d)):
        log(f"[{name}] unsupported"); return False
    rc = _lib.lib.octave_conv_tc_fwd(C.byref(d), dyn.data_ptr(), wp.data_ptr(), None, dx.data_ptr(), None, stream())
    torch.cuda.synchronize()
    if rc != 0:
        log(f"[{name}] rc={rc}"); return False
    return report(name, dx.permute(0, 3, 1, 2), ref)


def convt_case(name, B, H, W, cin, cout, Hout, Wout, ld_extra=0):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, cin, H, W, generator=g).bfloat16()
    w = (torch.randn(cin, cout, 2, 2, generator=g) / cin ** 0.5).bfloat16()
    b = torch.randn(cout, generator=g)
    ref = F.conv_transpose2d(x.float(), w.float(), b, stride=2)[:, :, :Hout, :Wout]
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    wp = w.permute(2, 3, 1, 0).reshape(1, 4 * cout, cin).contiguous().cuda()
    ld = cout + ld_extra
    y = torch.full((B, Hout, Wout, ld), float("nan"), dtype=torch.bfloat16, device="cuda")
    d = desc(B, H, W, cin, cout, 1, 1, y_ld=ld, mode=1, Hout=Hout, Wout=Wout)
    d.y_coff = ld_extra
    bc = b.cuda()
    rc = _lib.lib.octave_conv_tc_fwd(C.byref(d), xn.data_ptr(), wp.data_ptr(), bc.data_ptr(), y.data_ptr(), None, stream())
    torch.cuda.synchronize()
    if rc != 0:
        log(f"[{name}] rc={rc}"); return False
    return report(name, y[..., ld_extra:].permute(0, 3, 1, 2), ref)


def wgrad_case(name, B, H, W, cin, cout, groups, k):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, H, W, generator=g).bfloat16()
    dy = torch.randn(B, cout, H, W, generator=g).bfloat16()
    w = torch.zeros(cout, cin // groups, k, k, requires_grad=True)
    F.conv2d(x.float(), w, None, 1, k // 2, 1, groups).backward(dy.float())
    ref = w.grad   # the kernel writes the torch parameter layout [Cout][Cin/groups][k][k]
    xn = x.permute(0, 2, 3, 1).contiguous().cuda(); dyn = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dw = torch.full((cout, cin // groups, k, k), float("nan"), dtype=torch.float32, device="cuda")
    d = desc(B, H, W, cin, cout, groups, k)
    if not _lib.lib.octave_conv_tc_wgrad_supported(C.byref(d)):
        log(f"[{name}] unsupported"); return False
    rc = _lib.lib.octave_conv_tc_wgrad(C.byref(d), xn.data_ptr(), dyn.data_ptr(), dw.data_ptr(), stream())
    torch.cuda.synchronize()
    if rc != 0:
        log(f"[{name}] rc={rc}"); return False
    ok = report(name, dw, ref)
    if not ok:
        o = dw.cpu()
        log("   nan frac", torch.isnan(o).float().mean().item())
        e = (o - ref).abs().amax(dim=(0, 1)); log("   err per tap", e.tolist())
    return ok


def guarded(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception:
        log(f"[{a[0]}] EXCEPTION\n{traceback.format_exc()}")
        return False


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    log("device", torch.cuda.get_device_name(0), "sms", _lib.lib.octave_sm_count())
    res = {}
    if which in ("all", "fwd"):
        res["id1x1"] = guarded(conv_case, "id1x1_64", 1, 16, 16, 64, 64, 1, 1, bias=False, identity=True)
        res["1x1_64"] = guarded(conv_case, "1x1_64_64", 2, 20, 20, 64, 64, 1, 1)
        res["1x1_k256"] = guarded(conv_case, "1x1_256_128", 2, 25, 25, 256, 128, 1, 1)
        res["id3x3"] = guarded(conv_case, "id3x3_64", 1, 16, 16, 64, 64, 1, 3, bias=False, identity=True)
        res["3x3_64"] = guarded(conv_case, "3x3_64_64_40", 1, 40, 40, 64, 64, 1, 3)
        res["3x3_128"] = guarded(conv_case, "3x3_128_128_25", 2, 25, 25, 128, 128, 1, 3, relu=1)
        res["3x3_g2"] = guarded(conv_case, "3x3_256_256_g2_13", 2, 13, 13, 256, 256, 2, 3)
        res["3x3_bk32"] = guarded(conv_case, "3x3_64_128_g2_bk32", 2, 16, 16, 64, 128, 2, 3)
        res["3x3_32"] = guarded(conv_case, "3x3_32_64_bk32_n64", 1, 50, 50, 32, 64, 1, 3)
        res["1x1_n32"] = guarded(conv_case, "1x1_128_32", 1, 50, 50, 128, 32, 1, 1)
        res["3x3_big"] = guarded(conv_case, "3x3_512_256_100", 1, 100, 100, 512, 256, 1, 3)
    if which in ("all", "dgrad"):
        res["dgrad"] = guarded(dgrad_case, "dgrad_3x3_128_64_25", 2, 25, 25, 128, 64, 1, 3)
        res["dgrad_g"] = guarded(dgrad_case, "dgrad_3x3_g2", 2, 13, 13, 128, 256, 2, 3)
    if which in ("all", "convt"):
        res["convt"] = guarded(convt_case, "convt_128_64_13to25", 2, 13, 13, 128, 64, 25, 25)
        res["convt_cat"] = guarded(convt_case, "convt_64_64_concat", 1, 20, 20, 64, 64, 40, 40, ld_extra=64)
    if which in ("all", "wgrad"):
        res["wg1x1"] = guarded(wgrad_case, "wgrad_1x1_64_64", 2, 16, 16, 64, 64, 1, 1)
        res["wg3x3"] = guarded(wgrad_case, "wgrad_3x3_128_128_25", 2, 25, 25, 128, 128, 1, 3)
        res["wg3x3_256"] = guarded(wgrad_case, "wgrad_3x3_256_128_13", 2, 13, 13, 256, 128, 1, 3)
        res["wg3x3_g2"] = guarded(wgrad_case, "wgrad_3x3_128_256_g2", 2, 20, 20, 128, 256, 2, 3)
    log("SUMMARY", res)
