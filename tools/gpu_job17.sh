for ex in c5 c3,c5 c4,c5 north_star,c5; do
  timeout 600 python bench.py --no-cpu-baseline --extras $ex --steps 50 > /tmp/b.json 2>/tmp/b.err
  python - <<PY
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print('$ex', d['extras']['c5']['seconds'], d['extras']['c5']['ms_per_sequence_this_rank'])
PY
done
