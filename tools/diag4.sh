mkdir -p gpurun_out
for rep in 1 2; do
for v in old new; do
  for nh in 0 1; do
    if [ $v = old ]; then export QG_LIB=$PWD/tools/lib_old/libquivergpu.so; else unset QG_LIB; fi
    if [ $nh = 1 ]; then export QG_TC_NOHIT=1; else unset QG_TC_NOHIT; fi
    echo "== $v nohit=$nh"
    python tools/quickbench.py 2048 0,4 10 2>&1 | grep '"q"' | cut -c1-140
  done
done
done
unset QG_LIB; unset QG_TC_NOHIT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_ts_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_ts_lean python tools/prof_once.py 1000000 128 1 256 10 3 2>&1 | tail -1
