#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order.py tests/test_gpu_tester.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/e2e_profile.py 10 2>&1 | tail -9
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -3 gpurun_out/r2q_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2q_bench.json')); print(j['value'], j['ms_per_step'], j['e2e'], j['clocks'], j['roofline']['step_frac'])"
