mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scan.py tests/test_gpu_compact.py -x -q -m gpu 2>&1 | tail -3
for rk in 0 1; do
  export QG_TC_RAWK=$rk
  echo "== rawk $rk"
  python tools/quickbench.py 256,2048,10000 0,1,6 10 2>&1 | grep '"q"' | cut -c1-175
done
