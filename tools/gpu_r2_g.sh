#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:stem_pool_kernel" -s 1 -c 1 -o gpurun_out/r2g_stem_pool python tools/layer_report.py 256 > gpurun_out/r2g_ncu.log 2>&1; tail -1 gpurun_out/r2g_ncu.log
timeout 600 ncu --set full --clock-control none -k "regex:whdr_kernel|prf_kernel" -c 4 -o gpurun_out/r2g_metrics python -c "
import sys; sys.path.insert(0, '.')
import bench, torch
p = bench.measured_peaks()
print(bench.time_metrics('cuda:0', p, reps=1))" > gpurun_out/r2g_ncu_metrics.log 2>&1; tail -1 gpurun_out/r2g_ncu_metrics.log
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; cat gpurun_out/r2g_bench.json
