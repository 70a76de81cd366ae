echo "=== chain analysis at 4 x 300 (experiment build, carve-out default)"
for m in 0 1 2 4 8 5 13 3; do
  CFG=glamr_static_multi P=4 GLAMR_EXP_SKIP=$m GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done
echo "=== per-segment (eager) at 4 x 300"
CFG=glamr_static_multi P=4 timeout 300 python tools/iter_breakdown.py 2>&1 | tail -1
echo "=== launch list 4 x 300"
CFG=glamr_static_multi P=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lbs_|frame_residuals|traj_cam|pose_prep|apply_kernel|blend_features|camera_|cam_forward" -c 300 --csv --log-file gpurun_out/launches_r02i_p4.csv python tools/iter_breakdown.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02i_p4.csv | head -16
