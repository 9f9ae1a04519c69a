#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r26_trace.txt
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/r26_pytest.log
(timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r26_bench.json
cat gpurun_out/r26_trace.txt; tail -3 gpurun_out/r26_pytest.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r26_bench.json
