"""Opt-in per-entry-point GPU timing (CUDA events around every C-ABI call).  Development aid: not used on the hot path."""
from __future__ import annotations

import collections

import torch

from . import _lib

_records = []
_orig = {}


def _dims(name, args):
    """Structured shape record of a call (for the per-family rooflines of bench.py): conv descriptors and Act views."""
    d = {"name": name}
    if args and hasattr(args[0], "_obj"):
        o = args[0]._obj
        if isinstance(o, _lib.ConvDesc):
            d.update(kind="conv", B=o.B, H=o.H, W=o.W, Hout=o.Hout, Wout=o.Wout, cin=o.cin, cout=o.cout, k=o.ksize, g=o.groups, mode=o.mode)
        elif hasattr(o, "C") and hasattr(o, "H"):
            d.update(kind="act", B=o.B, H=o.H, W=o.W, C=o.C, esize=4 if o.dtype == _lib.DTYPE_F32 else 2)
            d["present"] = [i for i, a in enumerate(args) if a is not None and hasattr(a, "_obj")]
    return d


def _key(name, args):
    k = name
    if args and hasattr(args[0], "_obj"):
        o = args[0]._obj
        if isinstance(o, _lib.ConvDesc):
            k += f"[B{o.B} {o.H}x{o.W} {o.cin}->{o.cout} k{o.ksize} g{o.groups} m{o.mode}]"
        elif hasattr(o, "C") and hasattr(o, "H"):
            k += f"[{o.H}x{o.W}x{o.C}]"
    return k


def enable():
    if _orig:
        return
    for name in _lib.declared_symbols():
        fn = getattr(_lib.lib, name)
        if not name.endswith(("_fwd", "_bwd", "_wgrad", "_dgrad", "_stats", "_prepare", "_act", "_reduce", "_apply", "_inplace",
                              "_combine", "_du", "_nhwc", "_nchw", "_window", "_depth", "_space", "_weight", "_data", "_s2d",
                              "_noise", "_clipmask", "_multi")):
            continue
        _orig[name] = fn

        def wrap(*args, _fn=fn, _name=name):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = _fn(*args)
            e1.record()
            _records.append((_key(_name, args), e0, e1, _dims(_name, args)))
            return rc

        setattr(_lib.lib, name, wrap)


def disable():
    for name, fn in _orig.items():
        setattr(_lib.lib, name, fn)
    _orig.clear()


def reset():
    _records.clear()


def records():
    """[(key, milliseconds, dims)] of the recorded calls (synchronises)."""
    torch.cuda.synchronize()
    return [(k, e0.elapsed_time(e1), d) for k, e0, e1, d in _records]


def report(top: int = 40, by_shape: bool = False) -> str:
    torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, e0, e1, _d in _records:
        if not by_shape:
            k = k.split("[")[0]
        agg[k][0] += 1
        agg[k][1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    lines = [f"total {tot:.2f} ms over {len(_records)} calls"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% n={n:5d}  {k}")
    return "\n".join(lines)
