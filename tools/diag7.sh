mkdir -p gpurun_out
for v in old new; do
  for nh in 0 1; do
    if [ $v = old ]; then export QG_LIB=$PWD/tools/lib_old/libquivergpu.so; else unset QG_LIB; fi
    if [ $nh = 1 ]; then export QG_TC_NOHIT=1; else unset QG_TC_NOHIT; fi
    echo "== $v nohit=$nh"
    python tools/quickbench.py 2048 0,4 10 2>&1 | grep '"q"' | cut -c1-150
  done
done
