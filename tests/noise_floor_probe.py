"""TEST INFRASTRUCTURE (CPU): how far does rounding noise move the k-step optimisation trajectory?

    python tests/noise_floor_probe.py [case ...]

For every golden case prints, per tensor, max|ref32 - ref64| (the executed float32 reference against the float64
continuation stored in the fixture) next to max|emu - ref64| (the host-compiled frame functions of
glamr_b200/csrc/globalopt_frames.cuh + float32 Adam, joints from the oracle's SMPL).  The GPU parity test applies the
same yardstick to the CUDA path (tests/test_gpu_parity.py)."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from emu_runner import EmuRunner  # noqa: E402
from glamr_b200 import lib as L  # noqa: E402
from glamr_b200.synthetic import make_smpl_assets  # noqa: E402
from helpers import GLOBALOPT_CASES, ReplayMT, case_setup  # noqa: E402
from oracle.global_opt import OracleGlobalRecon  # noqa: E402


def emu_final_state(name, assets):
    gold, cfg, in_dict = case_setup(name, assets)
    ora = OracleGlobalRecon(cfg, assets, mt_model=ReplayMT(gold))
    data = ora.init_data(copy.deepcopy(in_dict))
    run = EmuRunner(ora, data)
    run.set_stage([], {}, 'init')
    run.backward()
    P, T = run.comp.P, run.comp.T
    losses = {}
    for stage, specs in cfg.opt_stage_specs.items():
        run.set_stage(specs['opt_variables'], specs['loss_cfg'], stage)
        for it in range(specs['opt_niters']):
            _, terms = run.backward()
            for k in specs['loss_cfg']:
                losses.setdefault(f'{stage}/{k}', []).append(float(terms[L.TERM_INDEX[k]]))
            run.step(specs['opt_lr'])
        cam = run.buffer(L.R_CAM_POSE).view(T, 3, 4)
        data['cam_pose'] = torch.cat([cam, torch.tensor([0., 0., 0., 1.]).expand(T, 1, 4)], dim=1).clone()
    out = {'cam_pose': data['cam_pose'].numpy()}
    ow, tw = run.buffer(L.R_ORIENT_WORLD).view(P, T, 3), run.buffer(L.R_TRANS_WORLD).view(P, T, 3)
    for p, pid in enumerate(data['person_data']):
        out[f'{pid}/smpl_orient_world'] = ow[p].numpy().copy()
        out[f'{pid}/root_trans_world'] = tw[p].numpy().copy()
        pv = run.layout.views(run.theta, p)
        for k in ['traj_local_xy', 'traj_local_rot', 'traj_local_z', 'traj_local_dxy']:
            out[f'{pid}/{k}'] = pv[k].numpy().copy()
    return gold, out, losses


def main(cases):
    assets = make_smpl_assets(0)
    for name in cases:
        gold, out, losses = emu_final_state(name, assets)
        print(f'== {name}')
        for k, v in out.items():
            if f'final64/{k}' not in gold:
                continue
            r64, r32 = gold[f'final64/{k}'], gold[f'final/{k}']
            rp = gold.get(f'final_pert/{k}')
            if k == 'cam_pose':
                r64, r32, v = r64[:, :3], r32[:, :3], v[:, :3]
            e_ref, e_emu = np.abs(r32 - r64).max(), np.abs(v.reshape(r64.shape) - r64).max()
            e_p = np.abs((rp[:, :3] if k == 'cam_pose' else rp) - r32).max() if rp is not None else 0.0
            print(f'  {k:28s} |ref32-ref64| {e_ref:9.2e}   |pert-ref32| {e_p:9.2e}   |emu-ref64| {e_emu:9.2e}   ratio {e_emu / max(e_ref, e_p, 1e-12):7.2f}')
        for k, v in losses.items():
            r64, r32 = gold[f'loss64/{k}'], gold[f'loss/{k}']
            sc = max(np.abs(r64).max(), 1e-12)
            print(f'  loss {k:40s} rel |ref32-ref64| {np.abs(r32 - r64).max() / sc:9.2e}   |emu-ref64| {np.abs(np.array(v) - r64).max() / sc:9.2e}')


if __name__ == '__main__':
    main(sys.argv[1:] or GLOBALOPT_CASES)
