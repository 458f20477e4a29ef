set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scan.py -x -q -m gpu 2>&1 | tail -5
python tools/quickbench.py 256,2048,10000 0,1,4,5 10 > gpurun_out/r02_quickbench_lean.txt 2>&1
grep '"q"' gpurun_out/r02_quickbench_lean.txt
