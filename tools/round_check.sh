# full GPU suite, the judged bench line, and the reference's configs at full size
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err | tail -3
