echo "=== gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench short"; timeout 600 python bench.py --no-cpu-baseline --extras north_star 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'e2e', d['e2e']['seconds'], {s:(round(x['ms_per_iter'],4)) for s,x in d['extras']['north_star']['stages'].items()})"
