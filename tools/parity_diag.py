"""Network-level parity diagnostics at the BASELINE shapes: CUDA path (bf16 / fp32 mode) against the CPU oracle (fp32, and
fp64 for the oracle's own noise floor).  Prints value errors, per-loss relative errors, whole-net gradient cosine.
  python tools/parity_diag.py <mode> <B> <H> [train|eval] [f64]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octave_b200 import config, network, losses, synth
from oracle import octave_oracle as O

mode, B, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
training = (sys.argv[4] if len(sys.argv) > 4 else "train") == "train"
want64 = len(sys.argv) > 5
torch.set_num_threads(os.cpu_count())


def l2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


config.set_compute_dtype(mode)
SEED = int(os.environ.get("DIAG_SEED", 0))
torch.manual_seed(SEED)
net = network.ResnestUNet(2, False)
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
net = net.cuda().train(training)
x, ys, _ = synth.octa_batch(B, H, H, seed=int(os.environ.get("DIAG_DATA_SEED", 7)))


def oracle(sd, x, ys):
    sdr = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    att, agg, x4 = O.segmentor_forward(sdr, x, training=training, st=O.BNState())
    wp = O.weighted_partial_ce(torch.softmax(agg, 1), ys, 2)
    kl = O.interlayer_divergence(att)
    loss = wp + 0.1 * kl
    names = [k for k, v in sdr.items() if v.requires_grad and not k.startswith("linear_head_")]
    gs = torch.autograd.grad(loss, [sdr[k] for k in names], allow_unused=True)
    return att, agg, wp, kl, dict(zip(names, gs))


t0 = time.time()
att_o, agg_o, wp_o, kl_o, g_o = oracle(sd, x, ys)
print(f"oracle fp32: {time.time() - t0:.1f} s")
if want64:
    t0 = time.time()
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    att_d, agg_d, wp_d, kl_d, g_d = oracle(sd64, x.double(), ys.double())
    print(f"oracle fp64: {time.time() - t0:.1f} s")
att, agg, x4 = net(x.cuda())
res = losses.FusedSegmentorLoss().total(agg, ys.cuda(), att, None, 1.0, 0.1, 0.0)
res['total'].backward()
params = dict(net.named_parameters())


def cos(ga, gb):
    num = da = db = 0.0
    worst = []
    for k, b in gb.items():
        a = ga[k] if isinstance(ga, dict) else None
        if b is None or a is None:
            continue
        a = a.detach().float().cpu().double().flatten(); b = b.double().flatten()
        num += float(a @ b); da += float(a @ a); db += float(b @ b)
        if float(b.norm()) > 0:
            worst.append((float((a - b).norm() / b.norm()), k))
    worst.sort(reverse=True)
    return num / (da ** 0.5 * db ** 0.5), (da / db) ** 0.5, worst


g_c = {k: p.grad for k, p in params.items() if p.grad is not None}
ref_name = "fp64 oracle" if want64 else "fp32 oracle"
A, G, WP, KL, GR = (att_d, agg_d, wp_d, kl_d, g_d) if want64 else (att_o, agg_o, wp_o, kl_o, g_o)
print(f"== {mode} B={B} {H}x{H} {'train' if training else 'eval'} vs {ref_name}")
print("agg l2", l2(agg, G), "att l2", [round(l2(a, b), 5) for a, b in zip(att, A)])
print("argmax equal:", bool(torch.equal(agg.argmax(1).cpu(), G.argmax(1))), "mismatch frac", float((agg.argmax(1).cpu() != G.argmax(1)).float().mean()))
print("wpce", float(res['supervised']), float(WP), "rel", abs(float(res['supervised']) - float(WP)) / abs(float(WP)))
print("kld ", float(res['divergence']), float(KL), "rel", abs(float(res['divergence']) - float(KL)) / abs(float(KL)))
c, r, w = cos(g_c, GR)
print("whole-net gradient cosine", c, "norm ratio", r)
print("worst per-parameter relative L2:", [(round(e, 4), k) for e, k in w[:6]], "median", round(w[len(w) // 2][0], 5))
# per top-level module: aggregated relative L2 error and cosine, in backward order
order = ["fc", "aag_0", "decoder_0", "upsampling_0", "aag_1", "decoder_1", "upsampling_1", "aag_2", "decoder_2", "upsampling_2", "aag_3",
         "decoder_3", "upsampling_3", "aag_4", "decoder_4", "upsampling_4", "encoder_4", "encoder_3", "encoder_2", "encoder_1", "encoder_0_1_2"]


def per_module(ga, gb, label):
    out = []
    for m in order:
        num = da = db = dd = 0.0
        for k, b in gb.items():
            if not k.startswith(m + ".") or b is None or ga.get(k) is None or k.endswith(("fc1.bias", "fc2.bias")):
                continue
            if ".conv" in k and k.endswith(".bias") and training:
                continue
            a = ga[k].detach().float().cpu().double().flatten(); b = b.double().flatten()
            num += float(a @ b); da += float(a @ a); db += float(b @ b); dd += float((a - b) @ (a - b))
        if db > 0:
            out.append(f"{m}:{(dd / db) ** 0.5:.3f}/{num / (da ** 0.5 * db ** 0.5 + 1e-300):.4f}")
    print(label, "relL2/cos per module:", " ".join(out))


per_module(g_c, GR, "cuda")
if "autocast" in sys.argv:
    # calibration: the SAME oracle arithmetic on stock torch CUDA ops under bf16 autocast (cuDNN), against the same reference
    dev = torch.device("cuda")
    sdr = {k: (v.clone().to(dev).requires_grad_() if v.is_floating_point() and "running" not in k else v.clone().to(dev)) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        att_a, agg_a, _ = O.segmentor_forward(sdr, x.to(dev), training=training, st=O.BNState())
    wp_a = O.weighted_partial_ce(torch.softmax(agg_a.float(), 1), ys.to(dev), 2)
    kl_a = O.interlayer_divergence([a.float() for a in att_a])
    names = [k for k, v in sdr.items() if v.requires_grad and not k.startswith("linear_head_")]
    gs = torch.autograd.grad(wp_a + 0.1 * kl_a, [sdr[k] for k in names], allow_unused=True)
    g_a = {k: g for k, g in zip(names, gs) if g is not None}
    c3, r3, w3 = cos(g_a, GR)
    print("autocast agg l2", l2(agg_a, G), "att l2", [round(l2(a, b), 5) for a, b in zip(att_a, A)], "agg absmax", float(G.abs().max()), "agg rms", float(G.pow(2).mean().sqrt()))
    print("torch bf16 autocast (cuDNN) vs", ref_name, ": cosine", c3, "norm ratio", r3, "median per-param", round(w3[len(w3) // 2][0], 5),
          "wpce rel", abs(float(wp_a) - float(WP)) / abs(float(WP)), "kld rel", abs(float(kl_a) - float(KL)) / abs(float(KL)), "agg l2", l2(agg_a, G))
    per_module(g_a, GR, "autocast")
if want64:
    per_module(g_o, g_d, "oracle32")
    c2, r2, w2 = cos(g_o, g_d)
    print("fp32 oracle vs fp64 oracle: cosine", c2, "agg l2", l2(agg_o, agg_d), "wpce rel", abs(float(wp_o) - float(wp_d)) / abs(float(wp_d)),
          "kld rel", abs(float(kl_o) - float(kl_d)) / abs(float(kl_d)), "median grad err", round(w2[len(w2) // 2][0], 6), "worst", round(w2[0][0], 5))
