timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity_chain.py tests/test_gpu_host.py tests/test_gpu_requests.py tests/test_gpu_filter_view.py tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -3
for c in 0; do QG_SCAN_TRACE=1 timeout 90 python tools/quickbench.py 1 $c 10 2>&1 | grep "scan trace: span\|scan trace: first\|scan trace: tau" | tail -3; done
timeout 300 python tools/quickbench.py 1,2,4,8 0,1,2,4,5,6 10,100 2>&1 | grep '"q"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if r['path'] != 1: continue
    print(r['n'], r['d'], r['metric'], 'q', r['q'], 'k', r['k'], 'ms', r['ms'], 'scan_us', r['prof']['scan_us'], 'fin_us', r['prof']['fin_us'], 'GB/s', r['gbs'], 'bad', r['bad'])"
