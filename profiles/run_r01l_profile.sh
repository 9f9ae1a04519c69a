#!/bin/bash
# (ncu passes run bench.py with E3DGE_BENCH_EAGER=1: kernels launched from Python in program order, so that the
# -s / -c windows below select the same launches as when these profiles were taken)
# Round-1 final-state measurement recipe (after the CTA-pair conv, TMA-store epilogues and the local feature query)
# Round-1 final-state measurement recipe (run under gpurun, one GPU): tests, bench (both arms), training-step and
# graph timings, launch lists, tile timeline, ncu --set full of the five hot kernels.  Numbers printed under ncu
# are never bench values.
mkdir -p gpurun_out
P=gpurun_out/r32
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > ${P}_pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -9) > ${P}_smoke.txt
(timeout 200 python profiles/time_local_query.py 2>&1 | tail -1) > ${P}_time_local_query.txt
(timeout 400 python bench.py --steps 20 --warmup 3 2>&1 | tail -1) > ${P}_bench.json
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1) > ${P}_bench_reference.json
(timeout 200 python profiles/time_backward.py 2>&1 | tail -6) > ${P}_time_backward.txt
(timeout 200 python profiles/time_graph.py 2>&1 | tail -2) > ${P}_time_graph.txt
(timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > ${P}_trace.txt
E3DGE_BENCH_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
E3DGE_BENCH_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file ${P}_launches_train.csv python profiles/time_backward.py 8 train_only > /dev/null 2>&1
E3DGE_BENCH_EAGER=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:siren_render_tc_kernel -s 3 -c 1 \
  -o ${P}_render_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_render.log 2>&1
E3DGE_BENCH_EAGER=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_kernel|tc_conv_pair_kernel" -s 9 -c 3 \
  -o ${P}_conv_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_conv.log 2>&1
E3DGE_BENCH_EAGER=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_upconv_phase_kernel|upconv_blur_act_kernel" -s 6 -c 4 \
  -o ${P}_upconv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_upconv.log 2>&1
E3DGE_BENCH_EAGER=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:siren_render_bwd_tc_kernel -s 1 -c 1 \
  -o ${P}_render_bwd_tc python profiles/time_backward.py 8 train_only > ${P}_ncu_render_bwd.log 2>&1
E3DGE_BENCH_EAGER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:local_query_kernel -s 5 -c 1 \
  -o ${P}_local_query python profiles/time_local_query.py > ${P}_ncu_local_query.log 2>&1
tail -3 ${P}_pytest.log; tail -3 ${P}_smoke.txt; cat ${P}_time_local_query.txt; cut -c1-400 ${P}_bench.json; cut -c1-300 ${P}_bench_reference.json; cat ${P}_time_backward.txt ${P}_time_graph.txt
