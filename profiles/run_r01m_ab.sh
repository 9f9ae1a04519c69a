#!/bin/bash
# CTA-pair render kernel v4 (direct relaxed remote arrives, head + tail split) vs default (EPI=3)
mkdir -p gpurun_out
(E3DGE_RENDER_EPI=7 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r20_pytest_epi7.log
(E3DGE_RENDER_EPI=7 timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r20_trace_epi7.txt
for epi in 3 7 3 7; do
  (E3DGE_RENDER_EPI=$epi timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> gpurun_out/r20_bench_epi$epi.json
done
tail -n 4 gpurun_out/r20_pytest_epi7.log; for f in gpurun_out/r20_bench_*.json; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f; grep -o '"kernel_ms": [0-9.]*' $f; done; cat gpurun_out/r20_trace_epi7.txt
