mkdir -p gpurun_out
echo "=== 1 smpl tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smpl" 2>&1 | tail -3
echo "=== 2 timing (experiment build): 8 epilogue warps in the skinning kernel"
GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
CFG=glamr_static_multi P=8 T=500 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
echo "=== 3 all gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== 4 launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|blend_features" -c 300 --csv --log-file gpurun_out/launches_r02j.csv python tools/iter_breakdown.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02j.csv | head -12
CFG=glamr_static_multi P=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|blend_features|camera_|cam_forward" -c 300 --csv --log-file gpurun_out/launches_r02j_p4.csv python tools/iter_breakdown.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02j_p4.csv | head -12
