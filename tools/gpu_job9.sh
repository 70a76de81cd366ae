mkdir -p gpurun_out
echo "=== 1 carve-out as a launch attribute (auto / 0 / 1), experiment build"
for c in auto 0 1; do
  GLAMR_SMEM_CARVEOUT=$c GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
  GLAMR_SMEM_CARVEOUT=$c CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done
GLAMR_SMEM_CARVEOUT=auto CFG=glamr_static_multi P=8 T=500 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
GLAMR_SMEM_CARVEOUT=0 CFG=glamr_static_multi P=8 T=500 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
GLAMR_SMEM_CARVEOUT=1 CFG=glamr_static_multi P=2 T=300 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
GLAMR_SMEM_CARVEOUT=0 CFG=glamr_static_multi P=2 T=300 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
echo "=== 2 gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== 3 bench"
timeout 900 python bench.py --no-cpu-baseline --extras north_star,c4 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; tail -2 gpurun_out/bench_r02h.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02h.json').read().strip().splitlines()[-1])
    print('ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'e2e', d['e2e']['seconds'], d['e2e'].get('phase_seconds'))
    for k,v in d['extras'].items(): print(k, {s:(round(x['ms_per_iter'],4), round(x['ms_per_iter_l2_warm'],4)) for s,x in v.get('stages',{}).items()}, v.get('e2e',{}).get('phase_seconds'))
except Exception as e: print('no json', e)
PY
