"""The C-ABI library loads on a CPU-only host and exports every symbol include/octave_b200.h declares."""
import ctypes

from octave_b200 import _lib


def test_every_declared_symbol_is_exported():
    syms = _lib.declared_symbols()
    assert "octave_loss_fwd" in syms and len(syms) >= 6
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/octave_b200.h but not exported"


def test_abi_version_and_struct_size():
    assert _lib.lib.octave_abi_version() == 4
    # OctaveLossDesc: 7 int32 + 5 + 5 int32 + 4 float + 3 float + 2 int32 + 1 float (jsd_eps) = 27 words
    assert ctypes.sizeof(_lib.LossDesc) == 27 * 4


def test_argument_validation_without_gpu():
    d = _lib.LossDesc()
    d.dtype, d.B, d.C, d.H, d.W, d.flags = 7, 1, 2, 16, 16, _lib.LOSS_WPCE
    assert _lib.lib.octave_loss_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == _lib.ERR_INVALID
    d.dtype, d.C = 0, 99
    assert _lib.lib.octave_loss_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == _lib.ERR_INVALID
    d.C = 3  # generic path is fp32 only
    d.dtype = 1
    assert _lib.lib.octave_loss_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == _lib.ERR_UNSUPPORTED
    d.dtype, d.C = 0, 2
    assert _lib.lib.octave_loss_uses_fast_path(ctypes.byref(d)) == 1
    d.H = 20
    assert _lib.lib.octave_loss_uses_fast_path(ctypes.byref(d)) == 0
    assert _lib.lib.octave_loss_stats_bytes(ctypes.byref(d)) == (32 + 2) * 8


def test_new_entry_points_validate_arguments_without_gpu():
    # narrow-layer conv variants answer "unsupported" for descriptors outside their shape class, never crash
    from octave_b200._lib import ConvDesc
    d = ConvDesc()
    d.B, d.H, d.W, d.cin, d.cout, d.groups, d.ksize, d.stride, d.pad = 1, 8, 8, 48, 48, 1, 3, 1, 1
    d.Hout, d.Wout, d.x_ld, d.y_ld = 8, 8, 48, 48
    assert _lib.lib.octave_conv_halo_supported(ctypes.byref(d)) == 0          # 48 channels: not a {32, 64} layer
    assert _lib.lib.octave_conv_halo_wgrad_supported(ctypes.byref(d)) == 0
    d.cin, d.cout, d.x_ld, d.y_ld = 64, 32, 64, 32
    d.in_dtype = d.out_dtype = _lib.DTYPE_BF16
    assert _lib.lib.octave_conv_halo_supported(ctypes.byref(d)) == 1
    assert _lib.lib.octave_pack_job_blocks(0, 32, 64, 1, 3) == (32 * 64 + 1023) // 1024
    assert _lib.lib.octave_pack_weight_multi(None, 0, 0, None) == _lib.ERR_INVALID


def test_integration_table_line_references_point_at_the_named_declarations():
    """INTEGRATION.md cites header line ranges for every C-ABI group: each entry point a row names (wildcards and a/b
    alternatives expanded) must be declared inside one of the ranges the row cites."""
    import fnmatch
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "octave_b200.h")).read().split("\n")
    decl = {}
    for i, line in enumerate(header):
        m = re.match(r"^(?:int|size_t|void|int64_t|unsigned long long|int32_t)\s+(octave_\w+)\(", line)
        if m:
            decl.setdefault(m.group(1), i + 1)
    rows = [r for r in open(os.path.join(root, "INTEGRATION.md")).read().split("\n") if r.startswith("| `octave_")]
    assert len(rows) >= 12
    checked = 0
    for row in rows:
        first = row.split("|")[1]
        cite = re.findall(r"\(`:([\d,\-]+)`\)", first)
        assert cite, first
        ranges = [tuple(int(x) for x in part.split("-")) for part in cite[-1].split(",")]
        names = re.findall(r"`(octave_[\w\*/\(\)]+)`", first)
        pats = []
        for n in names:
            if n.endswith("(_multi)"):                     # `octave_pack_weight(_multi)`: the name with and without the suffix
                pats += [n[:-8], n[:-8] + "_multi"]
                continue
            if "/" in n:                                   # `octave_loss_fwd/bwd`, `..._supported/_stats_bytes`
                base = n.split("/")[0]
                pats.append(base)
                for suf in n.split("/")[1:]:
                    pats.append(base.rsplit("_", 1)[0] + ("" if suf.startswith("_") else "_") + suf if not suf.startswith("_")
                                else base.rsplit("_", 1)[0] + suf)
            else:
                pats.append(n)
        for pat in pats:
            hits = [k for k in decl if fnmatch.fnmatch(k, pat)]
            for k in hits:
                assert any(a <= decl[k] <= b for a, b in ranges), (k, decl[k], ranges)
                checked += 1
    assert checked >= 60
