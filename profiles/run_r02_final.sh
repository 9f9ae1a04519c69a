#!/bin/bash
# Round-2 final-state recipe (one GPU, ~6 minutes under gpurun).  NOT run in this session after the last commits:
# the round's GPU budget ended with the scaling run (profiles/r02_scale.txt); the numbers in profiles/r02_* come
# from the interim runs named in profiles/README_r02.md.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
P=gpurun_out/r02f
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > ${P}_pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -9) > ${P}_smoke.txt
(timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1) > ${P}_bench.json
(timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > ${P}_bench_reference.json
(timeout 200 python profiles/time_local_branch.py 2>&1 | tail -6) > ${P}_time_local.txt
(timeout 200 python profiles/time_ops.py 2>&1 | tail -5) > ${P}_time_ops.txt
(timeout 300 python profiles/time_size1024.py 4 2>&1 | tail -2) > ${P}_time_1024.txt
E3DGE_BENCH_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-exact-fp32 \
  --no-local-branch --no-full-frame --no-size1024 > /dev/null 2>&1
E3DGE_BENCH_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file ${P}_launches_1024.csv python profiles/time_size1024.py 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -s 6 -c 6 \
  -o ${P}_tc_linear python profiles/time_local_branch.py 8 > ${P}_ncu_tc_linear.log 2>&1
E3DGE_BENCH_EAGER=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_kernel|tc_upconv_phase_kernel" \
  -s 12 -c 6 -o ${P}_narrow_conv python profiles/time_size1024.py 4 > ${P}_ncu_narrow.log 2>&1
tail -3 ${P}_pytest.log; tail -3 ${P}_smoke.txt; cut -c1-400 ${P}_bench.json; cat ${P}_time_local.txt ${P}_time_ops.txt ${P}_time_1024.txt
