#!/bin/bash
# parity-phase up-conv + render epilogue variants (EPI 0 / 1 / 3)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r16_pytest.log
for epi in 0 1 3; do
  (E3DGE_RENDER_EPI=$epi timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r16_bench_epi$epi.json
done
(E3DGE_RENDER_EPI=3 timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r16_trace_epi3.txt
(E3DGE_RENDER_EPI=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "generator or full or renderer" 2>&1 | tail -3) > gpurun_out/r16_pytest_epi3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r16_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -n 3 gpurun_out/r16_pytest.log gpurun_out/r16_pytest_epi3.log; for f in gpurun_out/r16_bench_*.json; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f; grep -o '"kernel_ms": [0-9.]*' $f | head -1; done; cat gpurun_out/r16_trace_epi3.txt
python profiles/summarize_ncu.py launches gpurun_out/r16_launches.csv 2>/dev/null | head -30
