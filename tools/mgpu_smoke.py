"""Multi-GPU smoke (torchrun): staged prints to locate hangs; persons sharded over ranks with one all-reduce per iteration."""
import copy, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glamr_b200 import lib as L
from glamr_b200.config import Config
from glamr_b200.recon import GlobalReconOptimizer
from glamr_b200.smpl import SMPL
from glamr_b200.synthetic import make_in_dict, make_smpl_assets, SyntheticPrior

def say(*a):
    print(f'[rank {os.environ.get("RANK")}] {time.time() % 1000:.1f}', *a, flush=True)

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
say('pg up')
x = torch.ones(4, device=dev); dist.all_reduce(x); torch.cuda.synchronize(); say('first allreduce ok', x[0].item())
a = make_smpl_assets(0)
cfg = Config('glamr_dynamic')
for st in cfg.opt_stage_specs.values(): st['opt_niters'] = 8
cfg.grecon_model_specs['use_cuda_graph'] = os.environ.get('GRAPH', '1') == '1'
in_dict = make_in_dict(a, world, 300)
m = GlobalReconOptimizer(cfg, dev, None, smpl=SMPL(a, device=dev), mt_model=SyntheticPrior(0, dev), dist=(rank, world))
say('model built')
out = m.optimize(copy.deepcopy(in_dict)); torch.cuda.synchronize()
say('optimize done', float(out['cam_pose'][0, 0, 0]), m.iter_ms)
if rank == 0:
    m1 = GlobalReconOptimizer(copy.deepcopy(cfg), dev, None, smpl=m.smpl, mt_model=SyntheticPrior(0, dev))
    ref = m1.optimize(copy.deepcopy(in_dict))
    import numpy as np
    say('vs single-GPU: max cam diff', float(np.abs(ref['cam_pose'] - out['cam_pose']).max()),
        'orient diff', max(float(np.abs(ref['person_data'][p]['smpl_orient_world'] - out['person_data'][p]['smpl_orient_world']).max()) for p in range(world)))
dist.barrier(); say('done'); dist.destroy_process_group()
