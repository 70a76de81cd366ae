mkdir -p gpurun_out
echo "=== 1 parity diagnostics (default path)"; timeout 300 python tools/parity_diag.py 3dpw_p2_t80_gaps static_multi_p4_t300 2>&1 | grep -v Warning | tee gpurun_out/parity_diag_r02b.txt | cut -c1-260
echo "=== 1b tests legacy+tc (pipelined blend)"; GLAMR_LBS_PATH=tc timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_legacy_tc.log 2>&1; tail -6 gpurun_out/t_legacy_tc.log
echo "=== 2 bench variants (cold / warm ms per step)"
for v in "legacy simt" "legacy tc" "fused tc"; do set -- $v; GLAMR_ITER_PATH=$1 GLAMR_LBS_PATH=$2 timeout 300 python bench.py --extras none --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2: cold',round(d['ms_per_step'],4),'warm',round(d['ms_per_step_l2_warm'],4),'lbs',round(d['roofline']['kernel_ms'],4),'e2e ms',round(d['e2e']['seconds']*1e3,1),d['e2e']['phase_seconds'])"; done
echo "=== 3 breakdown legacy+tc"; GLAMR_LBS_PATH=tc timeout 120 python tools/iter_breakdown.py 2>&1 | tail -1
echo "=== 4 prior variants"
GLAMR_PRIOR_GRAPH=1 GLAMR_NET_WIMG=1 timeout 300 python -m pytest tests -m gpu -q -k "prior or pipeline" 2>&1 | tail -3
for v in "0 0" "1 0" "0 1" "1 1"; do for shape in "64 120" "1 300"; do set -- $v $shape; GLAMR_PRIOR_GRAPH=$1 GLAMR_NET_WIMG=$2 B=$3 T=$4 timeout 120 python tools/prior_bench.py 2>&1 | tail -1 | sed "s/^/graph=$1 wimg=$2: /"; done; done
echo "=== 5 e2e breakdown C5-like (glamr_3dpw, gaps, YAML iterations)"; CFG=glamr_3dpw GAPS=1 timeout 200 python tools/e2e_breakdown.py 0 300 > gpurun_out/e2e_breakdown_c5.txt 2>&1; grep "wall ms" gpurun_out/e2e_breakdown_c5.txt | cut -c1-400; sed -n '/cumulative/,+22p' gpurun_out/e2e_breakdown_c5.txt | cut -c1-150
echo "=== 6 default bench, full"; BENCH_DEBUG=1 timeout 700 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -3 gpurun_out/bench_r02b.err; head -c 200 gpurun_out/bench_r02b.json; echo
echo "=== 7 memcheck (tc path)"; CASES=glamr_dynamic GLAMR_LBS_PATH=tc timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py > gpurun_out/memcheck_r02b.log 2>&1; tail -8 gpurun_out/memcheck_r02b.log
