#!/bin/bash
# ring-depth probes: forward 256-column tile with 2 vs 3 stages; multi-wave wide wgrad with one vs two CTAs per SM
for st in 3 2; do OCTAVE_FWD_STAGES=$st timeout 60 python tools/one_conv.py 32 100 512 256 3 1 fwd 2>&1 | tail -n 1 | sed "s/^/fwd stages=$st: /"; done
for st in 3 2; do OCTAVE_FWD_STAGES=$st timeout 60 python tools/one_conv.py 32 50 1024 512 3 1 fwd 2>&1 | tail -n 1 | sed "s/^/fwd stages=$st: /"; done
for two in 0 1; do OCTAVE_WGRAD_2CTA=$two timeout 60 python tools/one_conv.py 32 25 2048 1024 3 1 wgrad 2>&1 | tail -n 1 | sed "s/^/wgrad 2cta=$two: /"; done
for two in 0 1; do OCTAVE_WGRAD_2CTA=$two timeout 60 python tools/one_conv.py 32 25 1024 2048 3 1 wgrad 2>&1 | tail -n 1 | sed "s/^/wgrad 2cta=$two: /"; done
