#!/bin/bash
# Round-2 GPU pass A: full gpu test suite, sanitizer racecheck / synccheck on the pipeline kernels, ncu --set full of
# the gather and metric kernels, default bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python -m pytest tests/test_gpu_order.py -m gpu -q -s -k "realistic or benchmarked or match_reference" > gpurun_out/r2a_parity_numbers.log 2>&1
grep -E "head|256-pair|entries" gpurun_out/r2a_parity_numbers.log | tail -40
SEL='test_conv_bn_act and (B2_16x16_256-256_k3s1 or B2_64x64_64-64_k3s1 or B2_32x32_512-128_k1s1)'
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "test_conv_fused_pair or test_conv_dual" > gpurun_out/r2a_${tool}_fused.txt 2>&1
  tail -3 gpurun_out/r2a_${tool}_fused.txt
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "test_conv_bn_act" > gpurun_out/r2a_${tool}_conv.txt 2>&1
  tail -3 gpurun_out/r2a_${tool}_conv.txt
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_train_kernels.py -m gpu -q -x -k "test_conv_wgrad or test_bn_train" > gpurun_out/r2a_${tool}_train.txt 2>&1
  tail -3 gpurun_out/r2a_${tool}_train.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_patch_kernel -s 2 -c 2 -o gpurun_out/r2a_gather python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_gather.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:prf_kernel|whdr_kernel" -c 4 -o gpurun_out/r2a_metrics python -m pytest tests/test_gpu_metrics.py -m gpu -q > gpurun_out/r2a_ncu_metrics.log 2>&1
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
python tools/layer_report.py 256 > gpurun_out/r2a_layers.txt 2>&1
