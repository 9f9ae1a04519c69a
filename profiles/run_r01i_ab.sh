#!/bin/bash
# A/B call: banded up-conv (L2-resident G) and the packed-f32x2 render epilogue.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r15_pytest.log
for mb in 0 40 24 80; do kb=$((mb*1024));
  (E3DGE_UPCONV_BAND_KB=$kb timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260) > gpurun_out/r15_bench_band$mb.json
done
(E3DGE_RENDER_EPI=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r15_bench_epi1.json
(E3DGE_RENDER_EPI=0 timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r15_trace_epi0.txt
(E3DGE_RENDER_EPI=1 timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r15_trace_epi1.txt
(E3DGE_RENDER_EPI=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render or full or generator" 2>&1 | tail -3) > gpurun_out/r15_pytest_epi1.log
tail -3 gpurun_out/r15_pytest.log gpurun_out/r15_pytest_epi1.log; for f in gpurun_out/r15_bench_*.json; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f; grep -o '"kernel_ms": [0-9.]*' $f | head -1; done; cat gpurun_out/r15_trace_epi0.txt gpurun_out/r15_trace_epi1.txt
