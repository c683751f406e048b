#!/bin/bash
# pair heuristics + sub-sampled layer-end outputs: parity tests, layer report, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py tests/test_gpu_instadepth.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; tail -15 gpurun_out/r2i_tests.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2i_layers.txt 2>&1; head -42 gpurun_out/r2i_layers.txt
INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2i_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
INSTAORDER_SUBSAMPLE=0 INSTAORDER_PAIR=0 INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2i_bench_base.json 2> gpurun_out/r2i_bench_base.err; python -c "
import json; j=json.load(open('gpurun_out/r2i_bench_base.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['roofline']['step_frac'])"
