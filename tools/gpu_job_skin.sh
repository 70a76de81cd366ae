# 1-GPU job: tensor-core skinning (GLAMR_LBS_PATH=tc -> path 2) A/B against the default, init_data breakdown
mkdir -p gpurun_out
echo "=== 1 smpl tests (three paths)"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smpl" 2>&1 | tail -8
echo "=== 2 all gpu tests, GLAMR_LBS_PATH=tc"
GLAMR_LBS_PATH=tc timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== 3 bench A/B"
for p in tc tcblend; do
  GLAMR_LBS_PATH=$p timeout 600 python bench.py --no-cpu-baseline --extras none > gpurun_out/bench_skin_$p.json 2> gpurun_out/bench_skin_$p.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_skin_$p.json').read().strip().splitlines()[-1])
    print('$p', 'ms', d['ms_per_step'], 'warm', d['ms_per_step_l2_warm'], 'e2e', d['e2e']['seconds'], d['e2e'].get('phase_seconds'), 'parts', d['roofline'].get('kernel_parts',{}).get('critical_path_ms'), d['roofline'].get('tensor',{}) and d['roofline']['tensor'].get('kernel_ms'))
except Exception as e: print('no json', e)
PY
  tail -2 gpurun_out/bench_skin_$p.err | cut -c1-300
done
echo "=== 4 init_data breakdown"
timeout 300 python tools/init_breakdown.py 300 1 > gpurun_out/init_breakdown_p1.txt 2>&1; head -30 gpurun_out/init_breakdown_p1.txt
CFG=glamr_static_multi timeout 300 python tools/init_breakdown.py 300 4 > gpurun_out/init_breakdown_p4.txt 2>&1; head -16 gpurun_out/init_breakdown_p4.txt
echo "=== 5 ncu launch list + full on skin tc"
GLAMR_LBS_PATH=tc timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|blend_features" -c 400 --csv --log-file gpurun_out/launches_r02d.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/ncu_bench_d.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02d.csv | head -14
GLAMR_LBS_PATH=tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbs_skin_tc_kernel|pose_prep_kernel|lbs_blend_tc" --launch-skip 9 -c 3 -o gpurun_out/prof_skin_r02d -f python tools/iter_breakdown.py > gpurun_out/ncu_full_d.log 2>&1
tail -2 gpurun_out/ncu_full_d.log
echo "=== 6 memcheck path 2"
GLAMR_LBS_PATH=tc CASES=glamr_dynamic timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/memcheck_r02d.log 2>&1; tail -4 gpurun_out/memcheck_r02d.log
