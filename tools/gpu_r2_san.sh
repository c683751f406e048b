#!/bin/bash
# compute-sanitizer racecheck + synccheck + memcheck on the mbarrier / TMEM pipelines (conv, pair, row3, fused, dual tests)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -2
K='B3_64x64_64-64_k3s1_relu or B6_8x8_512-2048 or rows1280_c128_n64 or rows1000_c64_n64 or pair_B6 or pair_B3_32 or pair_B3_64x64_256'
for tool in racecheck synccheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "$K" > gpurun_out/r2_${tool}_conv.txt 2>&1; tail -4 gpurun_out/r2_${tool}_conv.txt
done
