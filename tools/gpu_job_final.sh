# end-of-round evidence run (1 GPU): tests, default bench (all legs), reference arm, smoke, ncu launch list + full capture, sanitizers
mkdir -p gpurun_out
echo "=== 1 gpu tests (defaults)"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/gpu_tests_final.log
echo "=== 2 smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== 3 bench (defaults)"; timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
    print('ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'launches', d['gpu_launches_per_step'], 'e2e', d['e2e']['seconds'], d['e2e'].get('phase_seconds'), 'clocks', d['clocks'])
    print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','kernel_ms','traffic')}, d['roofline'].get('kernel_parts',{}).get('critical_path_ms'), d['roofline'].get('tensor',{}).get('kernel_ms'), d['roofline'].get('tensor',{}).get('frac'))
    print('cpu', {k: d['cpu_baseline'].get(k) for k in ('value','cores','ms_per_iter_median','thread_sweep_ms_per_iter')})
    for k,v in d['extras'].items(): print(k, {s:(round(x['ms_per_iter'],4), round(x['ms_per_iter_l2_warm'],4)) for s,x in v.get('stages',{}).items()}, {kk:vv for kk,vv in v.items() if kk in ('ms_per_batch','seconds','sequences_per_sec')}, v.get('e2e',{}).get('phase_seconds'), v.get('e2e',{}).get('seconds'))
except Exception as e: print('no json', e)
PY
echo "=== 4 reference arm (short)"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref_final.err | tee gpurun_out/bench_ref_final.json | cut -c1-400
echo "=== 5 ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|blend_features" -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/ncu_bench_final.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02_final.csv | head -14
echo "=== 6 ncu --set full: skinning, blend, pose prep"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lbs_skin_tc_kernel|lbs_blend_tc_kernel|pose_prep_kernel" --launch-skip 12 -c 3 -o gpurun_out/prof_lbs_r02_final -f python tools/iter_breakdown.py > gpurun_out/ncu_full_final.log 2>&1; tail -2 gpurun_out/ncu_full_final.log | cut -c1-200
echo "=== 7 memcheck"; timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/memcheck_r02_final.log 2>&1; tail -5 gpurun_out/memcheck_r02_final.log
echo "=== 8 racecheck (glamr_dynamic)"; CASES=glamr_dynamic timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/racecheck_r02_final.log 2>&1; tail -3 gpurun_out/racecheck_r02_final.log; grep -c "Race reported" gpurun_out/racecheck_r02_final.log
