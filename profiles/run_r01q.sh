#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r25_trace.txt
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4) > gpurun_out/r25_pytest.log
(timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r25_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r25_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize_ncu.py launches gpurun_out/r25_launches.csv > gpurun_out/r25_launches.txt
cat gpurun_out/r25_trace.txt; tail -3 gpurun_out/r25_pytest.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r25_bench.json; sed -n 1,22p gpurun_out/r25_launches.txt
