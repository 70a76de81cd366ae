"""Host-side breakdown of GlobalReconOptimizer.optimize() (the e2e leg of bench.py): cProfile + wall-clock of a warm call."""
import copy, cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glamr_b200 import synthetic as syn
from glamr_b200.config import Config
from glamr_b200.recon import GlobalReconOptimizer
from glamr_b200.smpl import SMPL
from glamr_b200.motion_traj import MotionTrajJointModel
from glamr_b200.synthetic_nets import make_prior_states

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device('cuda:0')
assets = syn.make_smpl_assets(0)
cfg = Config(os.environ.get('CFG', 'glamr_dynamic'), out_dir='/tmp/e2e')
if K > 0:
    for st in cfg.opt_stage_specs.values():
        st['opt_niters'] = K
smpl = SMPL(assets, device=dev)
mt = MotionTrajJointModel(None, dev, None, smpl, make_prior_states())
model = GlobalReconOptimizer(cfg, dev, None, smpl=smpl, mt_model=mt)
P = int(sys.argv[3]) if len(sys.argv) > 3 else 1
in_dict = syn.make_in_dict(assets, P, T, seed=0, gaps=os.environ.get('GAPS', '0') == '1')
model.optimize(copy.deepcopy(in_dict))
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    model.optimize(copy.deepcopy(in_dict))
    torch.cuda.synchronize()
    print('optimize() wall ms', (time.perf_counter() - t0) * 1e3, 'K', K, 'phases', getattr(model, 'phase_seconds', None), 'iter_ms', getattr(model, 'iter_ms', None)[-2:])
pr = cProfile.Profile()
pr.enable()
model.optimize(copy.deepcopy(in_dict))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(40)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
