mkdir -p gpurun_out
P=gpurun_out/r2h
(timeout 300 python -m pytest tests/test_gpu_local_query.py -m gpu -q 2>&1 | tail -4) > ${P}_pytest.log
(timeout 300 python profiles/time_size1024.py 4 2>&1 | tail -2) > ${P}_time_1024.txt
E3DGE_BENCH_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file ${P}_launches_1024.csv python profiles/time_size1024.py 4 > /dev/null 2>&1
cat ${P}_pytest.log ${P}_time_1024.txt
