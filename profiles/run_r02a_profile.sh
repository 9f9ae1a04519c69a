#!/bin/bash
# Round-2 interim recipe (run under gpurun, one GPU): tests, stand-alone op roofline, launch list and ncu --set full
# of the local branch's tc_linear_kernel stages.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
P=gpurun_out/r2e
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > ${P}_pytest.log
(timeout 200 python profiles/time_ops.py 2>&1 | tail -5) > ${P}_time_ops.txt
(timeout 200 python profiles/time_local_branch.py 2>&1 | tail -6) > ${P}_time_local.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_linear_kernel|local_mlp_prep|siren_render_tc|local_query" -c 80 --csv \
  --log-file ${P}_launches_local.csv python profiles/time_local_branch.py 8 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear_kernel -s 6 -c 6 \
  -o ${P}_tc_linear python profiles/time_local_branch.py 8 > ${P}_ncu_tc_linear.log 2>&1
tail -3 ${P}_pytest.log; cat ${P}_time_ops.txt ${P}_time_local.txt
