# multi-GPU job: N=$1 (default 2); env knobs for the new paths are passed through
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for mode in nccl peer; do
  echo "=== smoke $mode"; GLAMR_ALLREDUCE=$mode timeout 300 $TR tools/mgpu_smoke.py 2>&1 | grep -E "vs single|optimize done|Error|error" | head -8
done
for mode in nccl peer; do
  echo "=== bench $mode N=$N"; GLAMR_ALLREDUCE=$mode timeout 600 $TR bench.py --gpus $N --steps 100 --extras north_star > gpurun_out/bench_mgpu_${mode}_n$N.json 2> gpurun_out/bench_mgpu_${mode}_n$N.err; tail -2 gpurun_out/bench_mgpu_${mode}_n$N.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_mgpu_${mode}_n$N.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'warm',d['ms_per_step_l2_warm'],'parity',d.get('parity',{}).get('max_abs'),'ok',d.get('parity',{}).get('ok'), 'north_star', {k:v['ms_per_iter'] for k,v in d['extras'].get('north_star',{}).get('stages',{}).items()})
except Exception as e: print('no json', e)
PY
done
