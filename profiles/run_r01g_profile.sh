#!/bin/bash
# Round-1 (session 3: backward kernels, trimmed epilogue) measurement recipe, run under gpurun.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/r8_pytest.log
(timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -3) > gpurun_out/r8_bench.log
(timeout 200 python profiles/time_backward.py 2>&1 | tail -8) > gpurun_out/r8_time_backward.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r8_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/r8_launches_train.csv python profiles/time_backward.py 2 train_only > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_tc_kernel -s 3 -c 1 \
  -o gpurun_out/r8_render_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_bwd_tc_kernel -s 1 -c 1 \
  -o gpurun_out/r8_render_bwd_tc python profiles/time_backward.py 8 train_only > gpurun_out/r8_ncu_render_bwd.log 2>&1
tail -8 gpurun_out/r8_pytest.log; cut -c1-400 gpurun_out/r8_bench.log; cat gpurun_out/r8_time_backward.txt
