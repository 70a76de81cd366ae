mkdir -p gpurun_out
echo "=== 1 iteration timing: blend early (two buffers) vs after skinning"
for e in 1 0; do
  GLAMR_BLEND_EARLY=$e GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=0 timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done
GLAMR_BLEND_EARLY=1 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=1 timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
for e in 1 0; do
  GLAMR_BLEND_EARLY=$e CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=0 timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done
echo "=== 2 gpu tests (defaults: early blend)"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== 3 bench"
timeout 900 python bench.py --no-cpu-baseline --extras north_star > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; tail -2 gpurun_out/bench_r02g.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02g.json').read().strip().splitlines()[-1])
    print('ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'e2e', d['e2e']['seconds'], d['e2e'].get('phase_seconds'))
    print('roofline', d['roofline'].get('kernel_parts'))
    for k,v in d['extras'].items(): print(k, {s:(round(x['ms_per_iter'],4), round(x['ms_per_iter_l2_warm'],4)) for s,x in v.get('stages',{}).items()}, v.get('e2e',{}).get('phase_seconds'))
except Exception as e: print('no json', e)
PY
