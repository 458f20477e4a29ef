mkdir -p gpurun_out
export QH_LIB=$PWD/tools/lib_tsan/libquiverhost.so
export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4 log_path=gpurun_out/r02_tsan"
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) timeout 900 python -m pytest tests/test_gpu_host.py -x -q -m gpu -k "concurrent or write_combined or delete_batch or update_batch" 2>&1 | tail -6 | tee gpurun_out/r02_tsan_pytest.txt
ls gpurun_out/r02_tsan* 2>/dev/null; for f in gpurun_out/r02_tsan.*; do grep -c "WARNING: ThreadSanitizer" $f; grep -A12 "WARNING: ThreadSanitizer" $f | head -60; done
