#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pair or banded or fused" 2>&1 | tail -6) > gpurun_out/r33_pytest.log
cat gpurun_out/r33_pytest.log
