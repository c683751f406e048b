#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_order.py -m gpu -x -q 2>&1 | tail -2
for v in "INSTAORDER_ST256=1" "INSTAORDER_ROW3=0"; do
  echo "== $v"
  env $v timeout 300 python tools/layer_report.py 256 2>&1 | grep -E "^pairs|^\s+(102|107|115|128|207|215|225|238|315|312) " | awk '{printf "%s:%s ", $1,$3}'; echo
done
