# full GPU suite, the judged bench line, the reference's configs at full size
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err | tail -2
timeout 1200 python tests/config_bench.py c1,c3,c4 1.0 > gpurun_out/configs_all.jsonl 2> gpurun_out/configs.err; tail -2 gpurun_out/configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_q10000_k10.csv python bench.py --steps 2 --warmup 1 --no-subrecords --no-cpu-baseline > /dev/null 2>&1; grep -c tc_ts gpurun_out/r02_launches_bench_q10000_k10.csv
