mkdir -p gpurun_out
echo "=== 1 gpu tests (defaults)"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== 2 bench (defaults, no cpu)"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; tail -3 gpurun_out/bench_r02e.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02e.json').read().strip().splitlines()[-1])
    print('ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'e2e', d['e2e']['seconds'], d['e2e'].get('phase_seconds'))
    print('roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel_ms','frac','achieved')}, d['roofline'].get('kernel_parts'), d['roofline'].get('tensor'))
    for k,v in d['extras'].items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','ms_per_batch','seconds','sequences_per_sec','value')}, v.get('e2e',{}).get('phase_seconds'))
except Exception as e: print('no json', e)
PY
echo "=== 3 init_data breakdown"
timeout 300 python tools/init_breakdown.py 300 1 > gpurun_out/init_breakdown_e_p1.txt 2>&1; head -34 gpurun_out/init_breakdown_e_p1.txt
echo "=== 4 frame sections (experiment build)"
GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/frame_sections.py 2>&1 | tail -16
GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so CFG=glamr_static_multi P=4 timeout 300 python tools/frame_sections.py 2>&1 | tail -16
echo "=== 5 iter breakdown"
timeout 300 python tools/iter_breakdown.py 2>&1 | tail -2
