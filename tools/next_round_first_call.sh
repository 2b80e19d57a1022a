#!/bin/bash
# First GPU call of the next round (~3 GPU-minutes): everything that was switched on at the very end of round 1 without a
# full-suite / bench re-run, plus the pending A/B of the 128-column slab kernel.
#   gpurun --timeout 600 -- tools/next_round_first_call.sh
set -o pipefail
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -n 6 | tee gpurun_out/r2_pytest.log
timeout 200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-600 gpurun_out/r2_bench.json
# 256-column kernel: previous 3-stage kernel for comparison (the default is the 4-stage slab kernel)
OCTAVE_FWD_STAGES=3 timeout 150 python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null | cut -c1-330 | sed 's/^/FWD_STAGES=3: /'
OCTAVE_WGRAD_2CTA=0 timeout 150 python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null | cut -c1-330 | sed 's/^/WGRAD_2CTA=0: /'
# 128-column kernel: 5 stages vs the unmeasured 6-stage slab variant
OCTAVE_FWD128_STAGES=5 timeout 60 python tools/slab_probe.py save /tmp/mid.pt mid 2>&1 | tail -n 4
OCTAVE_FWD128_STAGES=6 timeout 60 python tools/slab_probe.py compare /tmp/mid.pt mid 2>&1 | tail -n 8
OCTAVE_FWD128_STAGES=6 timeout 150 python bench.py --no-cpu --steps 5 --warmup 3 2>/dev/null | cut -c1-330 | sed 's/^/FWD128_STAGES=6: /'
