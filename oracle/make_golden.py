"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.pt by running the REAL reference (imported from
/root/reference, CPU fp32).  Run here (the reference cannot travel to the GPU box); the fixtures are
committed.  Usage:  python -m oracle.make_golden [losses|networks|all]
"""
from __future__ import annotations

import os
import sys

import torch

from oracle import refload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402


def _grad(loss, *ts):
    return [None if g is None else g.clone() for g in torch.autograd.grad(loss, ts, retain_graph=True, allow_unused=True)]


def make_losses(ref):
    cases = {}
    g = torch.Generator().manual_seed(11)
    for name, (B, C, H, W, L) in {"pyr32": (2, 2, 32, 32, 5), "pyr48x80": (3, 2, 48, 80, 5), "generic": (2, 3, 20, 28, 3)}.items():
        logits = torch.randn(B, C, H, W, generator=g)
        lab = torch.randint(0, C + 3, (B, H, W), generator=g)
        ys = torch.stack([(lab == c).float() for c in range(C)], 1)
        full = torch.stack([(lab % C == c).float() for c in range(C)], 1)
        if name == "generic":
            sizes = [(H, W), (H // 2, W // 2), (7, 9)]
        else:
            sizes = [(H >> k, W >> k) for k in range(L)]
        att = [torch.softmax(1.5 * torch.randn(B, C, h, w, generator=g), 1).requires_grad_() for h, w in sizes]
        logits.requires_grad_()
        yhat = torch.softmax(logits, 1)
        yhat_leaf = yhat.detach().clone().requires_grad_()
        d_real = torch.randn(B, 1, generator=g).requires_grad_()
        d_fake = torch.randn(B, 1, generator=g).requires_grad_()
        c = {"logits": logits.detach(), "yhat": yhat_leaf.detach(), "ys": ys, "full": full, "att": [a.detach() for a in att],
             "d_real": d_real.detach(), "d_fake": d_fake.detach()}
        wpce = ref.WeightedPartialCE(C, manual=True)
        l = wpce(yhat_leaf, ys.clone()); c["wpce"] = l.detach(); c["wpce_g"] = _grad(l, yhat_leaf)[0]
        l = wpce(yhat, ys.clone()); c["wpce_logits_g"] = _grad(l, logits)[0]
        l = wpce(yhat_leaf, ys.clone(), reduction="sum"); c["wpce_sum"] = l.detach()
        l = wpce(yhat_leaf, ys.clone(), full=True); c["wpce_full"] = l.detach(); c["wpce_full_g"] = _grad(l, yhat_leaf)[0]
        ysb = ys.clone(); l = wpce(yhat_leaf, ysb, ignore_bg=True); c["wpce_ignore_bg"] = l.detach(); c["ys_after_ignore_bg"] = ysb
        l = ref.DiceLoss()(yhat_leaf, full); c["dice"] = l.detach(); c["dice_g"] = _grad(l, yhat_leaf)[0]
        l = ref.InterlayerDivergence()(att); c["kld"] = l.detach(); c["kld_g"] = _grad(l, *att)
        w = [1.0, 2.0, 0.0, 0.5][: len(att) - 1]
        l = ref.InterlayerDivergence()(att, weights=w); c["kld_w"] = l.detach(); c["kld_w_g"] = _grad(l, *att)
        c["kld_weights"] = w
        l = ref.InterlayerDivergence(stop_gradient=True)(att); c["kld_stop_g"] = _grad(l, *att)
        l = ref.LSGeneratorLoss()(d_fake); c["lsg"] = l.detach(); c["lsg_g"] = _grad(l, d_fake)[0]
        l = ref.LSDiscriminatorialLoss()(d_real, d_fake); c["lsd"] = l.detach(); c["lsd_g"] = _grad(l, d_real, d_fake)
        cases[name] = c
    torch.save(cases, os.path.join(GOLD, "losses.pt"))
    print("losses.pt", {k: float(v["wpce"]) for k, v in cases.items()})


NET_SEED, NET_B, NET_H, NET_W = 123, 4, 64, 80          # 80/16 = 5 is odd: the pad/crop path of compose.py:125-147
BN_KEYS = ("encoder_0_1_2.1.running_mean", "encoder_2.0.conv2.bn0.running_var", "decoder_0.conv.1.running_mean",
           "decoder_3.downsample.1.running_var", "encoder_4.2.bn3.running_mean")


def make_networks(ref):
    """Seeded reference networks on one synthetic batch.  Only the seed travels: the host mirror's seeded construction
    is identical to the reference's (tests/test_host_mirror.py), so the consumer rebuilds the weights from NET_SEED."""
    out = {"seed": NET_SEED, "B": NET_B, "H": NET_H, "W": NET_W, "batch_seed": 21}
    x, ys, _ = synth.octa_batch(NET_B, NET_H, NET_W, seed=21)
    shape = (torch.Size((NET_B, 3, NET_H, NET_W)), torch.Size((NET_B, 2, NET_H, NET_W)))
    torch.manual_seed(NET_SEED)
    net = ref.OctaScribbleNet(*shape, True, False, instance_noise=False, label_noise=False)
    net.train()
    with torch.no_grad():
        att, agg, x4 = net.segmentor(x)
        sd = net.segmentor.state_dict()
        out["segmentor_train"] = {"att": [a.clone() for a in att], "agg": agg.clone(), "x4": x4.clone(),
                                  "bn": {k: sd[k].clone() for k in BN_KEYS}}
        pyr = synth.mask_pyramid(NET_B, NET_H, NET_W)
        logit = net.discriminator(pyr)
        dsd = net.discriminator.state_dict()
        out["discriminator_train"] = {"logit": logit.clone(),
                                      "u": {k: v.clone() for k, v in dsd.items() if k.endswith("weight_u")}}
        net.eval()
        att, agg, x4 = net.segmentor(x)
        out["segmentor_eval"] = {"att": [a.clone() for a in att], "agg": agg.clone(), "x4": x4.clone()}
    for name, gl in (("parallel_head", None), ("parallel_head_ag3", 3)):
        torch.manual_seed(NET_SEED)
        if gl is None:
            pn = ref.compose.ResnestUnetParallelHead(2, False)
        else:
            pn = ref.compose.ResnestUnetParallelHeadAttentionGate(2, False, None, gl)
        pn.train()
        with torch.no_grad():
            o = pn(x)
        if gl is None:
            out[name] = {"agg": o.clone()}
        else:
            (a, c), agg = o
            out[name] = {"agg": agg.clone(), "att": [t.clone() for t in a], "att_c": [t.clone() for t in c]}
    torch.save(out, os.path.join(GOLD, "networks.pt"))
    print("networks.pt", float(out["segmentor_train"]["agg"].abs().mean()), float(out["parallel_head"]["agg"].abs().mean()))


def main(which: str):
    ref = refload.load()
    if ref is None:
        raise SystemExit("/root/reference is not mounted")
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if which in ("losses", "all"):
        make_losses(ref)
    if which in ("networks", "all"):
        make_networks(ref)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
