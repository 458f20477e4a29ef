set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
export QG_LIB=$PWD/tools/lib_instr/libquivergpu.so
for cfg in "1000000 128 256 1" "1000000 96 256 1" "1000000 128 256 2" "1000000 128 256 0"; do
  echo "=== tc_timing $cfg" ; python tools/tc_timing.py $cfg 2>&1 | tail -12
done > gpurun_out/r02_tc_timing_base.txt 2>&1
unset QG_LIB
python tools/quickbench.py 256,2048,10000 0,1,4,5,6 10 > gpurun_out/r02_quickbench_base.txt 2>&1
tail -30 gpurun_out/r02_quickbench_base.txt
cat gpurun_out/r02_tc_timing_base.txt
