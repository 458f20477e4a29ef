# ncu evidence, round 2: the flat (dense) scan, and a single query on the bf16 stream
ncu --set full --clock-control none -k regex:scan_dense -s 3 -c 1 -o gpurun_out/r02_scan_dense_q1 -f python tools/prof_once.py 1000000 128 1 1 10 8 > gpurun_out/ncu_dense.log 2>&1; tail -1 gpurun_out/ncu_dense.log
ncu --set full --clock-control none -k regex:tc_ts_kernel -s 5 -c 1 -o gpurun_out/r02_tc_ts_q1 -f python tools/prof_once.py 1000000 128 1 1 10 8 > gpurun_out/ncu_tcq1.log 2>&1; tail -1 gpurun_out/ncu_tcq1.log
