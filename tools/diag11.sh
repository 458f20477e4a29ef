mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_chain.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; tail -3 gpurun_out/r02_bench_n1_a.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_n1_a.json'))
print({k: l[k] for k in ('value','ms_per_step','e2e','uncertified_queries_device_api','real_valued','small_batch_regime')})
print(l['roofline']); print(l['config']['parity_check'])"
