# ncu evidence for the flat (dense) scan: full-set capture of one launch + launch list of the Q=1 regime
ncu --set full --clock-control none --import-source on -k regex:scan_dense -s 3 -c 2 -o gpurun_out/r02_scan_dense_q1 -f python tools/prof_once.py 1000000 128 1 1 10 8 > gpurun_out/ncu_dense.log 2>&1; tail -2 gpurun_out/ncu_dense.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_flat_q1_k10.csv python tools/prof_once.py 1000000 128 1 1 10 12 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:scan_dense -s 3 -c 1 -o gpurun_out/r02_scan_dense_768_q1 -f python tools/prof_once.py 2000000 768 0 1 10 6 > gpurun_out/ncu_dense768.log 2>&1; tail -1 gpurun_out/ncu_dense768.log
ls -la gpurun_out/*.ncu-rep | tail -3
