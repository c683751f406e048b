#!/bin/bash
mkdir -p gpurun_out
K='B3_64x64_64-64_k3s1_relu or B6_8x8_512-2048 or rows1280_c128_n64 or rows1000_c64_n64 or pair_B6 or pair_B3_32 or pair_B3_64x64_256'
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "$K" > gpurun_out/r2_racecheck_conv.txt 2>&1; tail -4 gpurun_out/r2_racecheck_conv.txt
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "wgrad" > gpurun_out/r2_racecheck_wgrad.txt 2>&1; tail -4 gpurun_out/r2_racecheck_wgrad.txt
timeout 600 python tools/e2e_profile.py 10 > gpurun_out/r2_e2e_profile.txt 2>&1; cat gpurun_out/r2_e2e_profile.txt
python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2n_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'], j['cpu_baseline']['value'], j['training']['value'], j['training']['clocks'])"
