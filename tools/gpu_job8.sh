mkdir -p gpurun_out
echo "=== carve-out x blend placement (experiment build, lib graph)"
for c in 1 0; do for e in 0 1; do
  echo "carveout=$c early=$e"
  GLAMR_SMEM_CARVEOUT=$c GLAMR_BLEND_EARLY=$e GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=0 timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done; done
GLAMR_SMEM_CARVEOUT=1 GLAMR_BLEND_EARLY=0 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so GLAMR_EXP_SKIP=1 timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
for c in 1 0; do for e in 0 1; do
  echo "P=4 carveout=$c early=$e"
  GLAMR_SMEM_CARVEOUT=$c GLAMR_BLEND_EARLY=$e CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1
done; done
