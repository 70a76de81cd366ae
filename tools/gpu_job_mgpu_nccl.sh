# multi-GPU job, NCCL reduction only: N=$1, extras=$2 (default none)
N=${1:-2}; EX=${2:-none}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 100 --extras $EX > gpurun_out/bench_mgpu_nccl_n${N}_b.json 2> gpurun_out/bench_mgpu_nccl_n${N}_b.err; tail -2 gpurun_out/bench_mgpu_nccl_n${N}_b.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_mgpu_nccl_n${N}_b.json').read().strip().splitlines()[-1])
    print('N=$N value',d['value'],'ms',d['ms_per_step'],'warm',d['ms_per_step_l2_warm'],'parity',d.get('parity',{}).get('max_abs'),'ok',d.get('parity',{}).get('ok'), 'e2e', d['e2e']['seconds'])
    for k,v in d['extras'].items(): print(k, {s:(round(x['ms_per_iter'],4), round(x['ms_per_iter_l2_warm'],4)) for s,x in v.get('stages',{}).items()})
except Exception as e: print('no json', e)
PY
