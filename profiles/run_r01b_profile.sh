#!/bin/bash
# Round-1 (tensor-core path) measurement recipe, run under gpurun.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r5_pytest.log
(timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -3) > gpurun_out/r5_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r5_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_tc_kernel -s 3 -c 1 \
  -o gpurun_out/r5_render_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 15 -c 5 \
  -o gpurun_out/r5_conv_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5_ncu_conv.log 2>&1
tail -8 gpurun_out/r5_pytest.log; cut -c1-600 gpurun_out/r5_bench.log
