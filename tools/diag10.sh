mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for v in old new; do
    if [ $v = old ]; then export QG_LIB=$PWD/tools/lib_old/libquivergpu.so; else unset QG_LIB; fi
    echo "== $v"
    python tools/quickbench.py 2048,10000 0,1,4 10 2>&1 | grep '"q"' | cut -c1-200
done
done
