for pdl in 1; do for fw in 0 1 0 1; do
  export QG_FINALIZE_WARP=$fw QG_PDL=$pdl
  echo "== pdl $pdl finalize warp $fw"
  python tools/quickbench.py 10000 0 10,100 2>&1 | grep '"q"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['k'], r['ms'], r['prof'], 'bad', r['bad'], 'launches', r['launches'])"
done; done
