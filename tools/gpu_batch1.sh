#!/bin/bash
# tests + step profile + microbenches + ncu of two narrow convs
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/bench_ew.py > gpurun_out/bench_ew3.log 2>&1
python tools/profile_step.py 32 400 > gpurun_out/profile_step2.log 2>&1; head -40 gpurun_out/profile_step2.log
python tools/one_conv.py 32 100 64 256 1 > gpurun_out/one_conv_a.log 2>&1
python tools/one_conv.py 32 400 64 32 3 > gpurun_out/one_conv_b.log 2>&1
cat gpurun_out/one_conv_a.log gpurun_out/one_conv_b.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/conv_1x1_narrow -f python tools/one_conv.py 32 100 64 256 1 > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/conv_3x3_narrow -f python tools/one_conv.py 32 400 64 32 3 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
