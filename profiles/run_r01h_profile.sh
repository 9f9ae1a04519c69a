#!/bin/bash
# Round-1 final-state measurement recipe (run under gpurun): tests, bench, launch lists, ncu --set full of the
# three tensor-core kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r14_pytest.log
(timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1) > gpurun_out/r14_bench.json
(timeout 200 python profiles/time_backward.py 2>&1 | tail -6) > gpurun_out/r14_time_backward.txt
(timeout 200 python profiles/whatif_render.py 2>&1 | tail -8) > gpurun_out/r14_whatif.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r14_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/r14_launches_train.csv python profiles/time_backward.py 8 train_only > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_tc_kernel -s 3 -c 1 \
  -o gpurun_out/r14_render_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r14_ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_bwd_tc_kernel -s 1 -c 1 \
  -o gpurun_out/r14_render_bwd_tc python profiles/time_backward.py 8 train_only > gpurun_out/r14_ncu_render_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 15 -c 5 \
  -o gpurun_out/r14_conv_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r14_ncu_conv.log 2>&1
tail -3 gpurun_out/r14_pytest.log; cut -c1-300 gpurun_out/r14_bench.json; cat gpurun_out/r14_time_backward.txt
