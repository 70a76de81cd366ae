"""Shared test helpers: golden-fixture loading and a learned-prior stub that replays recorded outputs."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

GLOBALOPT_CASES = ['dynamic_p1_t40', 'static_p1_t24', 'static_multi_p3_t30', 'dynamic_multi_p2_t32', '3dpw_p2_t80_gaps', 'h36m_p1_t48_gaps']
# the shapes bench.py measures (+ a T > 512 track): reference goldens with more iterations, see tests/golden/make_golden.py
BENCH_SHAPE_CASES = ['dynamic_p1_t300', 'static_multi_p4_t300', 'static_multi_p2_t500', '3dpw_p1_t600_gaps']


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


class ReplayMT:
    """Stands in for MotionTrajJointModel: returns, call by call, what the reference's (seeded random-init)
    infiller + trajectory predictor produced when the fixture was generated."""

    def __init__(self, gold, device='cpu'):
        self.gold, self.calls, self.device = gold, 0, device

    supports_person_batch = True

    def inference(self, batch, sample_num=1):
        B = batch['in_body_pose'].shape[0]           # a batched call over persons consumes B consecutive recorded calls
        cat_dim = {'infer_out_body_pose': 0, 'infer_out_local_traj_tp': 1, 'infer_out_orient': 0, 'infer_out_trans': 0}
        out = {k: torch.cat([torch.tensor(self.gold[f'mt/{self.calls + b}/{k}'], device=self.device) for b in range(B)], dim=d)
               for k, d in cat_dim.items()}
        self.calls += B
        return out


def case_setup(name, smpl_assets):
    """-> (gold dict, Config with the fixture's iteration count, in_dict)"""
    from glamr_b200.config import Config
    from glamr_b200.synthetic import make_in_dict
    gold = load_golden('globalopt_' + name)
    P, T, gaps, niters = [int(v) for v in gold['meta']]
    cfg = Config(str(gold['cfg_id']))
    for st in cfg.opt_stage_specs.values():
        st['opt_niters'] = niters
    in_dict = make_in_dict(smpl_assets, P, T, seed=0, gaps=bool(gaps), seq_name=name)
    return gold, cfg, in_dict


C3_ROWS = [0, 5, 37, 63]


def c3_prior_inputs():
    """seeded inputs of the C3-shape prior run (BASELINE.json configs[2]: 64 sequences x 120 frames, frames 40-69
    invisible; one motion latent per window and one trajectory latent shared by the batch, injected through the reference's
    in_motion_latent / in_traj_latent).  Shared by tests/golden/make_golden.py (reference run) and the parity tests."""
    g = torch.Generator().manual_seed(64)
    pose = torch.randn(64, 120, 69, generator=g) * 0.3
    mask = torch.ones(64, 120)
    mask[:, 40:70] = 0
    return {'in_body_pose': pose * mask[..., None], 'frame_mask': mask, 'in_motion_latent': torch.randn(4, 128, generator=g),
            'in_traj_latent': torch.randn(1, 128, generator=g)}


from glamr_b200.synthetic import LatentInjector  # noqa: E402,F401  (re-exported for the tests)
