"""Section timing of frame_residuals_kernel (one warp, clock64 stamps; -DGLAMR_EXPERIMENT build only).

    python -c "from glamr_b200 import lib; lib.build_experiment()"          # here (cross-compiles)
    GLAMR_B200_SO=glamr_b200/libglamr_b200_exp.so python tools/frame_sections.py      # on the GPU box
"""
import copy, ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glamr_b200 import lib as L
from glamr_b200.config import Config
from glamr_b200.recon import GlobalReconOptimizer
from glamr_b200.smpl import SMPL
from glamr_b200.synthetic import make_in_dict, make_smpl_assets, SyntheticPrior
P, T = int(os.environ.get('P', 1)), int(os.environ.get('T', 300))
cfgid = os.environ.get('CFG', 'glamr_dynamic')
a = make_smpl_assets(0); dev = torch.device('cuda:0')
cfg = Config(cfgid); in_dict = make_in_dict(a, P, T)
m = GlobalReconOptimizer(cfg, dev, None, smpl=SMPL(a, device=dev), mt_model=SyntheticPrior(0, dev))
data = m.init_data(copy.deepcopy(in_dict))
stage, specs = list(cfg.opt_stage_specs.items())[-1]
m._cur_vars, m._cur_stage, m._loss_cfg = specs['opt_variables'], stage, specs['loss_cfg']
m._set_stage(data, specs['opt_variables'], specs['loss_cfg'], stage, reset_adam=True, begin=True)
hist = torch.zeros((400, L.NUM_TERMS + 1), device=dev)
lib = m._lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ['start->root/cam/Rs loaded', 'joint loop (raw_joint + kp terms)', 'warp sums', 'frame_rest entry', 'kp vjp', 'cam-frame pose + cam_traj', 'smoothness',
         'rel_transform', 'final aa vjp', 'stores', 'block term reduce']
acc = np.zeros((2, 11))
R = 20
for r in range(R + 3):
    flush.fill_(1)
    L.check(lib.glamr_opt_iterate(m._opt, L.ptr(m._theta), L.ptr(m._reduce), float(specs['opt_lr']), L.ptr(hist), L.NUM_TERMS + 1, 1, 0, L.stream_ptr()), 'iterate')
    out = (ctypes.c_longlong * 32)()
    L.check(lib.glamr_exp_frame_stamps(out), 'stamps')
    st = np.array(out[:]).reshape(2, 16)[:, :12]
    if r >= 3:
        acc += np.diff(st, axis=1)
acc /= R
print(f'P={P} T={T} {cfgid}:{stage}   cycles per section (CTA 0 | CTA 37), L2 flushed before every iteration')
for k, nme in enumerate(names):
    print(f'  {nme:40s} {acc[0, k]:9.0f} {acc[1, k]:9.0f}')
print(f'  {"total":40s} {acc[0].sum():9.0f} {acc[1].sum():9.0f}   (1965 cycles = 1 us)')
