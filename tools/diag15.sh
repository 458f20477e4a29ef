mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_q10000_k10.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-check --no-subrecords > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 2 -c 1 -f -o gpurun_out/r02_hnsw_search python tools/prof_hnsw.py 300000 128 10000 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none -k regex:finalize_cand_kernel -s 2 -c 1 -f -o gpurun_out/r02_finalize_cand python tools/prof_once.py 1000000 128 1 2048 10 3 2>&1 | tail -1
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_n1.json')); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['launch_ms'], l['roofline']['traffic'], l['small_batch_regime']['frac_of_measured_hbm'], l['clocks'])
print(json.load(open('gpurun_out/r02_bench_reference.json'))['value'])"
