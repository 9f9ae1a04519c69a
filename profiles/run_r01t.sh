#!/bin/bash
# CTA-pair conv kernel: parity, A/B against the single-CTA kernel on the same box, launch list
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r31_pytest.log
for v in 0 1 0 1; do (E3DGE_CONV_PAIR=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> gpurun_out/r31_bench_pair$v.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r31_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/summarize_ncu.py launches gpurun_out/r31_launches.csv > gpurun_out/r31_launches.txt
tail -4 gpurun_out/r31_pytest.log; for v in 0 1; do echo pair=$v; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r31_bench_pair$v.json; done; sed -n 3,12p gpurun_out/r31_launches.txt
