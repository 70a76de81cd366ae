"""cProfile of optimize() over DISTINCT sequences (the C5 sweep: every sequence has its own gaps -> its own parameter count -> a new
optimiser handle), to see what a new sequence costs beyond a repeated one."""
import copy, cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glamr_b200 import synthetic as syn
from glamr_b200.config import Config
from glamr_b200.recon import GlobalReconOptimizer
from glamr_b200.smpl import SMPL
from glamr_b200.motion_traj import MotionTrajJointModel
from glamr_b200.synthetic_nets import make_prior_states
dev = torch.device('cuda:0')
assets = syn.make_smpl_assets(0)
cfg = Config('glamr_3dpw', out_dir='/tmp/c5p')
smpl = SMPL(assets, device=dev)
mt = MotionTrajJointModel(None, dev, None, smpl, make_prior_states())
model = GlobalReconOptimizer(cfg, dev, None, smpl=smpl, mt_model=mt)
ins = [syn.make_in_dict(assets, 1, 300, seed=i, gaps=True, seq_name=f's_{i}') for i in range(10)]
model.optimize(copy.deepcopy(ins[0])); model.optimize(copy.deepcopy(ins[1]))
torch.cuda.synchronize()
for i in (2, 3, 3, 4):
    t0 = time.perf_counter(); model.optimize(copy.deepcopy(ins[i])); torch.cuda.synchronize()
    print(f'seq {i}: {1e3 * (time.perf_counter() - t0):.1f} ms  phases {model.phase_seconds}')
pr = cProfile.Profile(); pr.enable()
for i in (5, 6, 7):
    model.optimize(copy.deepcopy(ins[i]))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
