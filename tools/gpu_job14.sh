echo "=== features kernel at the top of the evaluation (experiment build timing)"
for f in 0 1; do
  GLAMR_FEATURES_EARLY=$f GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1 | sed "s/^/features_early=$f  /"
  GLAMR_FEATURES_EARLY=$f CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1 | sed "s/^/features_early=$f  /"
done
GLAMR_FEATURES_EARLY=1 CFG=glamr_static_multi P=8 T=500 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1 | sed "s/^/features_early=1  /"
echo "=== gpu tests with GLAMR_FEATURES_EARLY=1"
GLAMR_FEATURES_EARLY=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
