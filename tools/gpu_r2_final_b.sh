#!/bin/bash
# Round-2 final pass B (1 GPU, final build): full GPU test suite, layer report, default bench line, reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log; tail -3 gpurun_out/r2y_pytest.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2y_layers.txt 2>&1; head -12 gpurun_out/r2y_layers.txt
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2y_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['per_call']['value'], j['clocks'], j['roofline']['step_frac'], j['roofline']['frac'], j['cpu_baseline']['value'], j['cpu_baseline']['sample'][:40], j['training']['value'])"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err; cat gpurun_out/r2y_bench_reference.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
