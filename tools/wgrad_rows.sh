#!/bin/bash
# correctness + timing of the wide weight-gradient kernel for each stage geometry (OCTAVE_WGRAD_ROWS)
for r in ${ROWS_LIST:-128 64}; do
  echo "== rows $r"
  OCTAVE_WGRAD_ROWS=$r timeout 200 python -m pytest tests/test_conv_wgrad_gpu.py -x -q 2>&1 | tail -n 2
  for shp in "32 50 1024 512 3" "32 100 512 256 3" "32 25 2048 1024 3" "32 100 256 512 3" "32 100 256 64 1" "32 25 1024 256 1"; do
    OCTAVE_WGRAD_ROWS=$r timeout 100 python tools/one_conv.py $shp 1 wgrad 2>&1 | tail -n 1
  done
done
