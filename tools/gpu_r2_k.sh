#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; tail -3 gpurun_out/r2k_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2k_layers.txt 2>&1; head -12 gpurun_out/r2k_layers.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_fused_kernel|conv_row3_kernel" -s 14 -c 8 -o gpurun_out/r2k_fused python tools/layer_report.py 256 > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log
