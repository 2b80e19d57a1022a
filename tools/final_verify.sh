#!/bin/bash
# round-end verification: GPU test suite, default bench line, then the ncu launch list of one eager step
set -o pipefail
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 > gpurun_out/final_pytest.log
rc=$?
cat gpurun_out/final_pytest.log
grep -q "failed\|error" gpurun_out/final_pytest.log && { echo "TESTS FAILED - skipping the rest"; exit 1; }
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || { tail -n 20 gpurun_out/final_bench.err; exit 2; }
cat gpurun_out/final_bench.json
[ "$1" = "nolist" ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --launch-skip 1600 --launch-count 1500 \
  --log-file gpurun_out/launches_final2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/ncu_launch3.log 2>&1
wc -l gpurun_out/launches_final2.csv
