#!/bin/bash
# Round-2 final pass B (1 GPU, final build): parity tests of the kernels touched last, layer report, default bench line.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py tests/test_gpu_metrics.py tests/test_gpu_tester.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; tail -2 gpurun_out/r2y_pytest.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2y_layers.txt 2>&1; head -42 gpurun_out/r2y_layers.txt
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2y_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['per_call']['value'], j['clocks'], j['roofline']['step_frac'], j['roofline']['frac'], j['cpu_baseline']['value'], j['training']['value'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err; cat gpurun_out/r2y_bench_reference.json | cut -c1-600
