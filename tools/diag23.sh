timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for g in 1 0; do
  echo "== dense gather $g"
  QG_SCAN_DENSE_GATHER=$g timeout 900 python tests/config_bench.py c3 1.0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith('{'): continue
    r = json.loads(l)
    if 'ms' in r: print(r['config'][:70], 'q', r.get('q'), 'k', r.get('k'), 'ms', r['ms'], 'path', r.get('path'), 'GB/s', r.get('GBps'), r.get('parity'))"
done
