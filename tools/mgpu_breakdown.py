"""Per-kernel segment times of one multi-GPU iteration with the peer-memory reduction (torchrun; events between launches)."""
import copy, ctypes, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glamr_b200 import lib as L
from glamr_b200.config import Config
from glamr_b200.recon import GlobalReconOptimizer
from glamr_b200.smpl import SMPL
from glamr_b200.synthetic import make_in_dict, make_smpl_assets, SyntheticPrior
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
a = make_smpl_assets(0); cfg = Config('glamr_dynamic')
for st in cfg.opt_stage_specs.values(): st['opt_niters'] = 6
m = GlobalReconOptimizer(cfg, dev, None, smpl=SMPL(a, device=dev), mt_model=SyntheticPrior(0, dev), dist=(rank, world))
m.optimize(copy.deepcopy(make_in_dict(a, world, 300))); torch.cuda.synchronize()
lib = m._lib; stage, specs = list(cfg.opt_stage_specs.items())[-1]
hist = torch.zeros((4000, L.NUM_TERMS + 1), device=dev)
def it(n, graph):
    L.check(lib.glamr_opt_iterate(m._opt, L.ptr(m._theta), L.ptr(m._reduce), float(specs['opt_lr']), L.ptr(hist), L.NUM_TERMS + 1, n, graph, L.stream_ptr()), 'iterate')
it(5, 0)
L.check(lib.glamr_opt_kernel_timing(m._opt, 2), 't')
acc = None
for _ in range(40):
    dist.barrier(); it(1, 0); torch.cuda.synchronize()
    ms = (ctypes.c_float * 24)(); n = ctypes.c_int()
    L.check(lib.glamr_opt_kernel_times(m._opt, ms, ctypes.byref(n)), 'times')
    v = np.array(ms[:n.value]); acc = v if acc is None else acc + v
L.check(lib.glamr_opt_kernel_timing(m._opt, 0), 't')
print(f'[rank {rank}] peer_ok {m._peer_ok} eager segments us (traj_fwd, -, pose_prep, lbs, residuals, traj_bwd(+publish), apply(+wait)):', np.round(acc / 40 * 1000, 1).tolist(), flush=True)
# graph replays back to back
it(3, 1); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); it(200, 1); e1.record(); torch.cuda.synchronize()
print(f'[rank {rank}] graph loop {e0.elapsed_time(e1) / 200 * 1000:.1f} us / iteration', flush=True)
del m; torch.cuda.synchronize(); dist.barrier(); os._exit(0)
