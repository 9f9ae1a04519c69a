#!/bin/bash
# default = CTA pairs (EPI 7) with the small weights in shared memory; full GPU suite, A/B vs EPI 3, trace, launch list
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r22_pytest.log
(timeout 200 python profiles/trace_render.py 2>&1 | tail -16) > gpurun_out/r22_trace_epi7.txt
for epi in 3 7 3 7; do
  (E3DGE_RENDER_EPI=$epi timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> gpurun_out/r22_bench_epi$epi.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r22_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -n 4 gpurun_out/r22_pytest.log; for f in gpurun_out/r22_bench_*.json; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f; grep -o '"kernel_ms": [0-9.]*' $f; done; cat gpurun_out/r22_trace_epi7.txt
python profiles/summarize_ncu.py launches gpurun_out/r22_launches.csv | head -24
