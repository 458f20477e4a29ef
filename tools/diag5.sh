mkdir -p gpurun_out
export QG_TC_NOHIT=1
for m in 0 1 2; do
  export QG_TC_DBGMODE=$m
  echo "== dbgmode $m (nohit)"
  python tools/quickbench.py 2048 0,4 10 2>&1 | grep '"q"' | cut -c1-140
done
