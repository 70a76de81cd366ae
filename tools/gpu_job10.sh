echo "=== carve-out mask (1 GEMM kernels, 2 small SMPL kernels, 4 optimiser kernels), experiment build"
for m in 0 1 3 7 4 5; do
  echo "mask=$m"
  GLAMR_SMEM_CARVEOUT=$m GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
  GLAMR_SMEM_CARVEOUT=$m CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done
