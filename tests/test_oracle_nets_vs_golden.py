"""Pin the oracle's learned-prior restatement (oracle/nets.py) against outputs of the executed reference networks
(tests/golden/nets.npz: reference modules with the seeded stand-in weights, injected latents).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from glamr_b200.synthetic_nets import make_prior_states
from oracle.nets import MotionTrajJoint
from oracle.smpl import OracleSMPL

NETS_CASES = ['b3_t75', 'b1_t300', 'b2_t40']


@pytest.fixture(scope='module')
def joint_model(smpl_assets):
    st_m, st_t = make_prior_states(1234)
    return MotionTrajJoint(st_m, st_t, OracleSMPL(smpl_assets))


@pytest.mark.parametrize('tag', NETS_CASES)
def test_motion_traj_inference_matches_reference(tag, joint_model):
    g = load_golden('nets')
    batch = {k: torch.tensor(g[f'{tag}/in/{k}']) for k in ['in_body_pose', 'frame_mask', 'in_motion_latent', 'in_traj_latent']}
    out = joint_model.inference(batch, sample_num=1)
    for k, tol in [('infer_out_body_pose', 2e-5), ('infer_out_local_traj_tp', 2e-5), ('infer_out_orient', 1e-4), ('infer_out_trans', 1e-4),
                   ('infer_out_pose', 1e-4)]:
        np.testing.assert_allclose(out[k].numpy(), g[f'{tag}/{k}'], atol=tol, err_msg=f'{tag} {k}')


def test_c3_shape_matches_reference(joint_model):
    """BASELINE.json configs[2] shape: 64 x 120 with frames 40-69 masked"""
    from helpers import C3_ROWS, c3_prior_inputs
    g = load_golden('nets')
    out = joint_model.inference(c3_prior_inputs(), sample_num=1)
    sel = torch.tensor(C3_ROWS)
    for k, bdim, tol in [('infer_out_body_pose', 0, 2e-5), ('infer_out_local_traj_tp', 1, 2e-5), ('infer_out_trans', 0, 1e-4), ('infer_out_orient', 0, 1e-4)]:
        np.testing.assert_allclose(out[k].index_select(bdim, sel).numpy(), g[f'c3_b64_t120/{k}'], atol=tol, err_msg=k)
        red = [d for d in range(out[k].dim()) if d != bdim]
        np.testing.assert_allclose(out[k].double().abs().sum(dim=red).numpy(), g[f'c3_b64_t120/{k}/abs_sum'], rtol=1e-5, err_msg=k)
