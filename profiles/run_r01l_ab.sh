#!/bin/bash
# CTA-pair render kernel with local publishes + forwarder (EPI=7) vs default (EPI=3); ncu of the blur kernel
mkdir -p gpurun_out
(E3DGE_RENDER_EPI=7 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "generator_vs_reference or full_size or point_queries" 2>&1 | tail -5) > gpurun_out/r19_pytest_epi7.log
(E3DGE_RENDER_EPI=7 timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r19_trace_epi7.txt
for epi in 3 7; do
  (E3DGE_RENDER_EPI=$epi timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r19_bench_epi$epi.json
done
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r19_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r19_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -n 4 gpurun_out/r19_pytest_epi7.log gpurun_out/r19_pytest.log; for f in gpurun_out/r19_bench_*.json; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f; grep -o '"kernel_ms": [0-9.]*' $f | head -1; done; cat gpurun_out/r19_trace_epi7.txt
python profiles/summarize_ncu.py launches gpurun_out/r19_launches.csv 2>/dev/null | grep -i "blur\|phase\|conv_kernel"
