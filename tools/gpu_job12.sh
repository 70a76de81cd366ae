echo "=== split blend: share launched at the top of the evaluation (experiment build for timing)"
for sp in 0 30 60; do
  GLAMR_BLEND_SPLIT=$sp GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1 | sed "s/^/split=$sp  /"
done
for sp in 0 20 30 50; do
  GLAMR_BLEND_SPLIT=$sp CFG=glamr_static_multi P=4 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1 | sed "s/^/split=$sp  /"
done
for sp in 0 15 30; do
  GLAMR_BLEND_SPLIT=$sp CFG=glamr_static_multi P=8 T=500 GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so timeout 300 python tools/iter_skip_exp.py 2>&1 | tail -1 | sed "s/^/split=$sp  /"
done
echo "=== gpu tests with GLAMR_BLEND_SPLIT=30"
GLAMR_BLEND_SPLIT=30 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
