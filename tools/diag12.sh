mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_a.json 2> gpurun_out/r02_bench_n2_a.err; tail -5 gpurun_out/r02_bench_n2_a.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_n2_a.json'))
print({k: l[k] for k in ('value','ms_per_step','e2e','uncertified_queries_device_api')})
print(json.dumps(l['row_sharded'], indent=1)); print(l['config']['parallelism'])"
