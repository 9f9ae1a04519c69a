#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r29_pytest.log
for i in 1 2; do (timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> gpurun_out/r29_bench.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r29_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize_ncu.py launches gpurun_out/r29_launches.csv > gpurun_out/r29_launches.txt
tail -4 gpurun_out/r29_pytest.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r29_bench.json; sed -n 3,14p gpurun_out/r29_launches.txt
