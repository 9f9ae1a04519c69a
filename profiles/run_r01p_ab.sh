#!/bin/bash
# fused up-conv -> conv pair (inference) + graph-vs-eager step timing
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r24_pytest.log
for i in 1 2; do
  (timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> gpurun_out/r24_bench.json
done
(timeout 200 python profiles/time_graph.py 2>&1 | tail -2) > gpurun_out/r24_time_graph.txt
tail -n 6 gpurun_out/r24_pytest.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r24_bench.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r24_bench.json; cat gpurun_out/r24_time_graph.txt
