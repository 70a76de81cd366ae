mkdir -p gpurun_out
echo "=== 1 defaults (legacy iteration, SIMT LBS; SMPL tests cover both LBS paths)"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t1_default.log 2>&1; tail -25 gpurun_out/t1_default.log
echo "=== 2 fused+simt"; GLAMR_ITER_PATH=fused timeout 600 python -m pytest tests -m gpu -q -k "globalopt or residual or graph or optimize or pipeline" > gpurun_out/t2_fused_simt.log 2>&1; tail -15 gpurun_out/t2_fused_simt.log
echo "=== 3 fused+tc+prior graphs+wimg"; GLAMR_ITER_PATH=fused GLAMR_LBS_PATH=tc GLAMR_PRIOR_GRAPH=1 GLAMR_NET_WIMG=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t3_new.log 2>&1; tail -25 gpurun_out/t3_new.log
echo "=== 4 bench default"; BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; tail -3 gpurun_out/bench_r02a.err; head -c 300 gpurun_out/bench_r02a.json; echo
echo "=== 5 bench fused+simt"; GLAMR_ITER_PATH=fused timeout 300 python bench.py --extras none --no-cpu-baseline > gpurun_out/bench_r02a_fused_simt.json 2>/dev/null; head -c 300 gpurun_out/bench_r02a_fused_simt.json; echo
echo "=== 6 bench fused+tc+graphs+wimg"; GLAMR_ITER_PATH=fused GLAMR_LBS_PATH=tc GLAMR_PRIOR_GRAPH=1 GLAMR_NET_WIMG=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02a_new.json 2>/dev/null; head -c 300 gpurun_out/bench_r02a_new.json; echo
echo "=== 6b prior variants (ms per batch: B=64 T=120 then B=1 T=300)"
for v in "0 0" "1 0" "0 1" "1 1"; do set -- $v; for shape in "64 120" "1 300"; do set -- $v $shape; GLAMR_PRIOR_GRAPH=$1 GLAMR_NET_WIMG=$2 B=$3 T=$4 timeout 120 python tools/prior_bench.py 2>&1 | tail -1 | sed "s/^/graph=$1 wimg=$2: /"; done; done
echo "=== 7 breakdowns"
timeout 120 python tools/iter_breakdown.py 2>&1 | tail -1
GLAMR_ITER_PATH=fused timeout 120 python tools/iter_breakdown.py 2>&1 | tail -1
GLAMR_ITER_PATH=fused GLAMR_LBS_PATH=tc timeout 120 python tools/iter_breakdown.py 2>&1 | tail -1
GLAMR_ITER_PATH=fused GLAMR_LBS_PATH=tc P=4 CFG=glamr_static_multi timeout 120 python tools/iter_breakdown.py 2>&1 | tail -1
timeout 200 python tools/e2e_breakdown.py 200 300 > gpurun_out/e2e_breakdown_r02a.txt 2>&1; grep "wall ms" gpurun_out/e2e_breakdown_r02a.txt
GLAMR_ITER_PATH=fused GLAMR_LBS_PATH=tc GLAMR_PRIOR_GRAPH=1 timeout 200 python tools/e2e_breakdown.py 200 300 > gpurun_out/e2e_breakdown_r02a_new.txt 2>&1; grep "wall ms" gpurun_out/e2e_breakdown_r02a_new.txt
echo "=== 8 ncu launch list"; GLAMR_ITER_PATH=fused GLAMR_LBS_PATH=tc timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|forward_pose|residuals_backward|cam_forward|camera_" -c 300 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 12 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | head -c 300; echo
