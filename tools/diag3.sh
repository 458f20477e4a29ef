set -x
mkdir -p gpurun_out
for v in old new; do
  if [ $v = old ]; then export QG_LIB=$PWD/tools/lib_old/libquivergpu.so; else unset QG_LIB; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_ts_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_ts_$v python tools/prof_once.py 1000000 128 1 256 10 3 2>&1 | tail -3
done
ls -la gpurun_out/*.ncu-rep
