mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r35_pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4) > gpurun_out/r35_smoke.txt
(timeout 400 python bench.py --steps 20 --warmup 3 2>&1 | tail -1) > gpurun_out/r35_bench.json
tail -3 gpurun_out/r35_pytest.log; tail -2 gpurun_out/r35_smoke.txt; cut -c1-330 gpurun_out/r35_bench.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r35_bench.json
