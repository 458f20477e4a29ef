mkdir -p gpurun_out
export QG_LIB=$PWD/tools/lib_instr/libquivergpu.so
QG_TC_TRACE=gpurun_out/r02_trace_l2_hits.txt python tools/tc_timing.py 1000000 128 256 1 2>&1 | tail -3
QG_TC_NOHIT=1 QG_TC_TRACE=gpurun_out/r02_trace_l2_nohit.txt python tools/tc_timing.py 1000000 128 256 1 2>&1 | tail -1
QG_TC_TRACE=gpurun_out/r02_trace_dot_hits.txt python tools/tc_timing.py 1000000 128 256 2 2>&1 | tail -1
wc -l gpurun_out/r02_trace_*.txt
