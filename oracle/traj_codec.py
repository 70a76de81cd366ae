"""ORACLE (test infrastructure): the 11-D local <-> global trajectory codec, restated from
traj_pred/utils/traj_utils.py (reference).  Layout per frame:
[dx, dy (heading frame; frame 0 = absolute xy), z, local orient 6d, cos/sin of d_heading (frame 0 = absolute)].
"""
import numpy as np
import torch
from scipy.interpolate import interp1d

from . import rotations as rt

BASE_ORIENT = (0.5, 0.5, 0.5, 0.5)


def rot_2d(xy, theta):
    """traj_utils.py:7-11"""
    c, s = torch.cos(theta), torch.sin(theta)
    return torch.stack([xy[..., 0] * c - xy[..., 1] * s, xy[..., 0] * s + xy[..., 1] * c], dim=-1)


def local_to_global(local_traj, local_heading=True):
    """traj_utils.py:65-88 (local_orient_type='6d', deheading_local=False) -> trans [T,...,3], orient_q [T,...,4].
    Time is dim 0."""
    base = torch.tensor(BASE_ORIENT, dtype=local_traj.dtype, device=local_traj.device)
    d_xy_h, z = local_traj[..., :2], local_traj[..., 2]
    d6, hvec = local_traj[..., 3:-2], local_traj[..., -2:]
    d_heading = rt.vec_to_heading(hvec)
    heading = torch.cumsum(d_heading, dim=0) if local_heading else d_heading
    heading_q = rt.heading_to_quat(heading)
    d_xy = torch.cat([d_xy_h[:1], rot_2d(d_xy_h[1:], heading[:-1])], dim=0)
    xy = torch.cumsum(d_xy, dim=0)
    trans = torch.cat([xy, z.unsqueeze(-1)], dim=-1)
    q = rt.quat_mul(heading_q, rt.rot6d_to_quat(d6))
    q = rt.quat_mul(q, base.expand_as(q))
    return trans, q


def global_to_local(trans, orient_q):
    """traj_utils.py:44-62 (local_orient_type='6d')"""
    base = torch.tensor(BASE_ORIENT, dtype=trans.dtype, device=trans.device)
    xy, z = trans[..., :2], trans[..., 2]
    q = rt.quat_mul(orient_q, rt.quat_conj(base).expand_as(orient_q))
    heading = rt.get_heading(q)
    local_q = rt.deheading_quat(q, rt.get_heading_q(q))
    d6 = rt.quat_to_rot6d(local_q)
    d_heading = torch.cat([heading[:1], heading[1:] - heading[:-1]])
    hvec = rt.heading_to_vec(d_heading)
    d_xy = torch.cat([xy[:1], rot_2d(xy[1:] - xy[:-1], -heading[:-1])])
    return torch.cat([d_xy, z.unsqueeze(-1), d6, hvec], dim=-1)


def _lin_interp(vis_ind, values, n):
    f = interp1d(vis_ind, values, axis=0, assume_sorted=True, fill_value='extrapolate')
    return f(np.arange(n, dtype=np.float32))


def interp_orient_q_sep_heading(orient_q_vis, vis_frames):
    """traj_utils.py:120-142: interpolate heading vector and heading-free 6d orientation separately over
    invisible frames (SciPy, linear, extrapolating), then recompose."""
    base = torch.tensor(BASE_ORIENT, dtype=orient_q_vis.dtype)
    q = rt.quat_mul(orient_q_vis, rt.quat_conj(base).expand_as(orient_q_vis))
    hq = rt.get_heading_q(q)
    hvec = rt.heading_to_vec(rt.get_heading(q))
    d6 = rt.quat_to_rot6d(rt.deheading_quat(q, hq))
    n = vis_frames.shape[0]
    vis_ind = torch.where(vis_frames)[0].numpy()
    hvec_i = torch.tensor(_lin_interp(vis_ind, hvec.numpy(), n), dtype=torch.float32)
    d6_i = torch.tensor(_lin_interp(vis_ind, d6.numpy(), n), dtype=torch.float32)
    out = rt.quat_mul(rt.heading_to_quat(rt.vec_to_heading(hvec_i)), rt.rot6d_to_quat(d6_i))
    return rt.quat_mul(out, base.expand_as(out))
