#!/bin/bash
# Round-1 measurement recipe (run under gpurun): parity suite, ncu launch list, ncu full capture.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2_pytest.log
# launch list: every kernel launch of a short bench with its device time (cold, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r2_bench_under_ncu.log 2>&1
# full capture of the dominant kernels (one launch each, after warm-up)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:siren_render_kernel -s 3 -c 1 \
  -o gpurun_out/r2_render python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_ffma_kernel -s 8 -c 2 \
  -o gpurun_out/r2_conv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_conv.log 2>&1
# CPU port: which thread count is fastest on this host?
timeout 600 python - > gpurun_out/r2_cpu_threads.log 2>&1 <<'PY'
import os, sys, time, torch
sys.path[:0] = ['.', 'tests']
from helpers import synthetic_state_dict, decoder_layout
from oracle import params as P, stylesdf_oracle as O
sd = synthetic_state_dict(256, 64, 2024, 'sharp')
inp = P.make_inputs(2024, 1, decoder_layout(256, 64), 64)
print('cpu_count', os.cpu_count())
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        with torch.no_grad():
            O.generator_forward(sd, inp['w'], inp['w_dec'], inp['cam_poses'], inp['focal'], inp['near'], inp['far'])
        ts.append(time.perf_counter() - t0)
    print(nt, 'threads', min(ts[1:]), 's/frame')
PY
tail -15 gpurun_out/r2_pytest.log; cat gpurun_out/r2_cpu_threads.log | tail -8; ls -la gpurun_out
