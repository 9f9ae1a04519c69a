#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r34_pytest.log
for i in 1 2; do (timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> gpurun_out/r34_bench.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r34_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize_ncu.py launches gpurun_out/r34_launches.csv > gpurun_out/r34_launches.txt
tail -3 gpurun_out/r34_pytest.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r34_bench.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r34_bench.json; sed -n 3,14p gpurun_out/r34_launches.txt
