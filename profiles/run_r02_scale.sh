#!/bin/bash
# Round-2 scaling recipe (run under gpurun --gpus 8): weak (8 frames per GPU) and strong (global batch 32,
# BASELINE.json configs[2]) scaling of the generator pass over 1 / 2 / 4 / 8 GPUs of one box.
mkdir -p gpurun_out
P=gpurun_out/r2s
X="--steps 10 --warmup 3 --no-cpu-baseline --no-exact-fp32 --no-local-branch --no-full-frame --no-size1024"
for n in 1 2 4 8; do
  for mode in weak strong; do
    extra=""; [ $mode = strong ] && extra="--scaling strong --global-batch 32"
    if [ $n = 1 ]; then
      timeout 300 python bench.py --gpus 1 $X $extra 2> ${P}_${mode}_${n}.err | tail -1 > ${P}_${mode}_${n}.json
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port $((29500 + n)) bench.py --gpus $n $X $extra 2> ${P}_${mode}_${n}.err | tail -1 > ${P}_${mode}_${n}.json
    fi
    python - <<PY
import json
try:
    d = json.loads(open("${P}_${mode}_${n}.json").read())
    print("$mode", $n, "GPUs:", round(d["value"], 1), "frames/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["value"], 1), "|", d["config"]["launch"][:90])
except Exception as e:
    print("$mode", $n, "FAILED", e)
PY
  done
done
