mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_driver.py > gpurun_out/r02_memcheck.log 2>&1; tail -4 gpurun_out/r02_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_driver.py > gpurun_out/r02_racecheck.log 2>&1; tail -4 gpurun_out/r02_racecheck.log
