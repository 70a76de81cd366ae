# multi-GPU job: N=$1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for mode in nccl peer; do
  echo "=== bench $mode N=$N"; GLAMR_ALLREDUCE=$mode timeout 600 $TR bench.py --gpus $N --steps 100 --extras north_star,c4 > gpurun_out/bench_mgpu_${mode}_n$N.json 2> gpurun_out/bench_mgpu_${mode}_n$N.err; tail -2 gpurun_out/bench_mgpu_${mode}_n$N.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_mgpu_${mode}_n$N.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'warm',d['ms_per_step_l2_warm'],'parity',d.get('parity',{}).get('max_abs'),'ok',d.get('parity',{}).get('ok'))
    print('where', d.get('parity',{}).get('where',{}).get('kp_2d_pred'))
    for k,v in d['extras'].items(): print(k, {s:(round(x['ms_per_iter'],4), round(x['ms_per_iter_l2_warm'],4)) for s,x in v.get('stages',{}).items()})
except Exception as e: print('no json', e)
PY
done
