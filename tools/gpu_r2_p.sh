#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/layer_report.py 256 2>&1 | grep -E "^pairs|^\s+(107|115|128|207|215|225|238|315|325|335|345|312) " | awk '{printf "%s:%s ", $1,$3}'; echo
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2p_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
