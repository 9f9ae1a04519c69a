#!/bin/bash
# Round-1 (pipelined tensor-core renderer, persistent conv) measurement recipe, run under gpurun.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r7_pytest.log
(timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -3) > gpurun_out/r7_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r7_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_tc_kernel -s 3 -c 1 \
  -o gpurun_out/r7_render_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r7_ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 15 -c 5 \
  -o gpurun_out/r7_conv_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r7_ncu_conv.log 2>&1
tail -8 gpurun_out/r7_pytest.log; cut -c1-300 gpurun_out/r7_bench.log
