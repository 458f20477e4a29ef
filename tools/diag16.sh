mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_scan.py tests/test_gpu_filter_view.py tests/test_gpu_parity_chain.py tests/test_gpu_host.py tests/test_shard_merge.py -x -q -m gpu 2>&1 | tail -4
for fw in 0 1 0 1; do
  export QG_FINALIZE_WARP=$fw
  echo "== finalize warp $fw"
  python tools/quickbench.py 2048,10000 0,1,4 10,100 2>&1 | grep '"q"' | cut -c1-215
done
