#!/bin/bash
# Round-2 final pass A (1 GPU): full GPU test suite, layer report, ncu launch list + --set full captures (summarised on
# the box: the .ncu-rep files are too large to travel), bench lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -4 gpurun_out/r2z_pytest.log
timeout 300 python tools/layer_report.py 256 > gpurun_out/r2z_layers.txt 2>&1; head -3 gpurun_out/r2z_layers.txt
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2z_bench.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['per_call']['value'], j['clocks'], j['roofline']['step_frac'], j['cpu_baseline']['value'], j['training']['value'], j['roofline']['metrics'])"
INSTAORDER_BENCH_TRAIN=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:conv_row3|conv_fused|conv_tc_kernel|stem_pool|conv_tn" -s 38 -c 38 -o /tmp/r2z_net python tools/layer_report.py 256 > gpurun_out/r2z_ncu_net.log 2>&1; tail -1 gpurun_out/r2z_ncu_net.log
python tools/ncu_summary.py /tmp/r2z_net.ncu-rep > gpurun_out/r2z_ncu_net.md 2>&1; head -5 gpurun_out/r2z_ncu_net.md
INSTAORDER_BENCH_TRAIN=0 timeout 900 ncu --set full --clock-control none -k "regex:gather_patch|prf_kernel|whdr_kernel|decide_kernel|tail_kernel" -s 6 -c 8 -o /tmp/r2z_io python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_ncu_io.log 2>&1
python tools/ncu_summary.py /tmp/r2z_io.ncu-rep > gpurun_out/r2z_ncu_io.md 2>&1; cat gpurun_out/r2z_ncu_io.md
for m in resize384 instadepth384 c1_o256 c3_ordernet256; do
  INSTAORDER_BENCH_TRAIN=0 timeout 600 python bench.py --mode $m --steps 20 --no-cpu-baseline > gpurun_out/r2z_bench_$m.json 2> gpurun_out/r2z_bench_$m.err; python -c "
import json,sys; j=json.load(open('gpurun_out/r2z_bench_$m.json')); print('$m', j['value'], j['ms_per_step'], j['e2e']['value'])"
done
timeout 600 python bench.py --workload train --steps 30 > gpurun_out/r2z_train.json 2> gpurun_out/r2z_train.err; python -c "
import json; j=json.load(open('gpurun_out/r2z_train.json')); print('train', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['step_frac'])"
timeout 600 python tools/train_report.py 32 > gpurun_out/r2z_train_report.txt 2>&1; head -9 gpurun_out/r2z_train_report.txt
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
