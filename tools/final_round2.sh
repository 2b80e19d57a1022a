#!/bin/bash
# Round-2 evidence in one GPU call: GPU test suite, bench lines of c2 / c4 / c5, loss-kernel table, smoke twice, parity log.
#   gpurun --timeout 1500 -- tools/final_round2.sh
set -o pipefail
mkdir -p gpurun_out/final
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -n 4 > gpurun_out/final/pytest_gpu.log; cat gpurun_out/final/pytest_gpu.log
for c in c2 c4 c5; do
  timeout 400 python bench.py --config $c > gpurun_out/final/bench_${c}.json 2> gpurun_out/final/bench_${c}.err || tail -n 5 gpurun_out/final/bench_${c}.err
  cut -c1-260 gpurun_out/final/bench_${c}.json
done
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final/bench_reference_c2.json 2>/dev/null; cut -c1-300 gpurun_out/final/bench_reference_c2.json
timeout 120 python tools/bench_loss.py > gpurun_out/final/loss_k9.log 2>&1; cat gpurun_out/final/loss_k9.log
bash tools/smoke_twice.sh 2>&1 | tee gpurun_out/final/smoke_twice.log
{
  for m in bf16 fp32; do timeout 300 python tools/parity_diag.py $m 2 304 train f64 autocast 2>&1 | grep -v Warning | grep -v "^  print" | tail -n 12; done
  timeout 400 python tools/parity_diag.py bf16 8 400 train autocast 2>&1 | grep -v Warning | grep -v "^  print" | tail -n 10
  timeout 300 python tools/parity_diag.py bf16 8 304 eval 2>&1 | grep -v Warning | grep -v "^  print" | tail -n 8
} > gpurun_out/final/parity.log 2>&1
tail -n 4 gpurun_out/final/parity.log
