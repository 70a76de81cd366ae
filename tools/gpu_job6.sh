mkdir -p gpurun_out
echo "=== 1 gpu tests (defaults; skin sweeps 3 frame tiles per CTA)"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== 2 which chain bounds the iteration (experiment build)"
for mask in 0 1 2 4 8 5 9 13 3; do
  GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=$mask timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done
echo "=== 3 bench"
timeout 900 python bench.py --no-cpu-baseline --extras north_star > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; tail -2 gpurun_out/bench_r02f.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02f.json').read().strip().splitlines()[-1])
    print('ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'e2e', d['e2e']['seconds'], d['e2e'].get('phase_seconds'))
    print('roofline', d['roofline'].get('kernel_parts'), d['roofline'].get('tensor'))
    for k,v in d['extras'].items(): print(k, {s:(round(x['ms_per_iter'],4), round(x['ms_per_iter_l2_warm'],4)) for s,x in v.get('stages',{}).items()}, v.get('e2e',{}).get('phase_seconds'))
except Exception as e: print('no json', e)
PY
echo "=== 4 launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|blend_features" -c 400 --csv --log-file gpurun_out/launches_r02f.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/ncu_bench_f.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02f.csv | head -14
