mkdir -p gpurun_out
P=gpurun_out/r2i
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > ${P}_pytest.log
(timeout 300 python profiles/time_size1024.py 4 2>&1 | tail -2) > ${P}_time_1024.txt
tail -8 ${P}_pytest.log; cat ${P}_time_1024.txt
