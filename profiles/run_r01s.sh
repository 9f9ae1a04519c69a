#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_local_query.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r30_pytest_local.log
(timeout 200 python profiles/time_local_query.py 2>&1 | tail -1) > gpurun_out/r30_time_local_query.txt
cat gpurun_out/r30_pytest_local.log gpurun_out/r30_time_local_query.txt
