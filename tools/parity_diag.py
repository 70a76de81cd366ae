"""Diagnostic (GPU): full per-tensor report of the k-step comparison for golden cases -- |cuda - ref64| next to |ref32 - ref64|,
position of the largest deviation, init-state deviations, per-iteration loss deviations.  Test infrastructure only."""
import copy
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from glamr_b200 import lib as L  # noqa: E402
from glamr_b200.recon import GlobalReconOptimizer  # noqa: E402
from glamr_b200.synthetic import make_smpl_assets  # noqa: E402
from helpers import ReplayMT, case_setup  # noqa: E402

DEV = 'cuda:0'
a = make_smpl_assets(0)
for name in sys.argv[1:] or ['3dpw_p2_t80_gaps', 'static_multi_p4_t300']:
    gold, cfg, in_dict = case_setup(name, a)
    model = GlobalReconOptimizer(cfg, torch.device(DEV), None, smpl=a, mt_model=ReplayMT(gold, DEV))
    data = model.init_data(copy.deepcopy(in_dict))
    print(f'== {name}')
    for pid, pd in data['person_data'].items():
        for k in ['kp_2d_pred', 'smpl_orient_world', 'root_trans_world', 'traj_local_pred', 'smpl_pose', 'smpl_orient_cam', 'root_trans_cam', 'person2cam']:
            if f'init/{pid}/{k}' in gold:
                d = np.abs(pd[k].cpu().numpy().astype(np.float64) - gold[f'init/{pid}/{k}'])
                print(f'  init {pid}/{k:20s} max|diff| {d.max():.3e}')
    print(f'  init cam_pose max|diff| {np.abs(data["cam_pose"].cpu().numpy() - gold["init/cam_pose"]).max():.3e}')
    for stage, specs in cfg.opt_stage_specs.items():
        n = specs['opt_niters']
        model.optimize_main(data, specs['opt_variables'], specs['opt_lr'], n, specs['loss_cfg'], {'stage': stage})
        hist = model.loss_history.cpu().numpy()
        for k in specs['loss_cfg']:
            r32, r64 = gold[f'loss/{stage}/{k}'], gold[f'loss64/{stage}/{k}']
            got = hist[:n, L.TERM_INDEX[k]]
            sc = max(np.abs(r64).max(), 1e-12)
            print(f'  loss {stage}/{k:30s} it0 rel {abs(got[0] - r64[0]) / sc:.2e}  max rel |cuda-ref64| {np.abs(got - r64).max() / sc:.2e}  |ref32-ref64| {np.abs(r32 - r64).max() / sc:.2e}')
    checks = [('cam_pose', data['cam_pose'].cpu().numpy())]
    for pid, pd in data['person_data'].items():
        for k in ['smpl_orient_world', 'root_trans_world', 'traj_local_xy', 'traj_local_dxy', 'traj_local_z', 'traj_local_rot', 'traj_local_heading',
                  'traj_local_dheading', 'world_dheading', 'kp_2d_pred']:
            if k in pd and f'final64/{pid}/{k}' in gold:
                checks.append((f'{pid}/{k}', pd[k].cpu().numpy()))
    for key, got in checks:
        r32, r64 = gold[f'final/{key}'], gold[f'final64/{key}']
        g = got.reshape(r64.shape)
        e = np.abs(g - r64)
        pos = np.unravel_index(np.argmax(e), e.shape)
        print(f'  final {key:24s} |cuda-ref64| {e.max():.3e} at {pos} (ref64 {r64[pos]:.6f} ref32 {r32[pos]:.6f} cuda {g[pos]:.6f})  |ref32-ref64| {np.abs(r32 - r64).max():.3e}  |cuda-ref32| {np.abs(g - r32).max():.3e}')
