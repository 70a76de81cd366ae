"""Rotation / rigid-transform helpers for the host-side (one-off) init_data mirror.

Every non-trivial formula runs in the CUDA library through the row-wise entry points (glamr_rowop_fwd), i.e. the
same device functions the fused optimisation kernels use; torch is only used for reshapes, concatenation and small
matmuls on the device tensors.  Names follow the reference modules lib/utils/torch_transform.py and
lib/utils/konia_transform.py so call sites read like the reference.
"""
import torch

from . import lib as L


def angle_axis_to_rotation_matrix(aa):
    return L.rowop(L.ROP_AA_TO_ROTMAT, aa).reshape(aa.shape[:-1] + (3, 3))


def rotation_matrix_to_quaternion(R):
    return L.rowop(L.ROP_ROTMAT_TO_QUAT, R.reshape(R.shape[:-2] + (9,)))


def quaternion_to_angle_axis(q):
    return L.rowop(L.ROP_QUAT_TO_AA, q)


def angle_axis_to_quaternion(aa):
    return L.rowop(L.ROP_AA_TO_QUAT, aa)


def quaternion_to_rotation_matrix(q):
    return L.rowop(L.ROP_QUAT_TO_ROTMAT, q).reshape(q.shape[:-1] + (3, 3))


def rotation_matrix_to_angle_axis(R):
    return L.rowop(L.ROP_ROTMAT_TO_AA, R.reshape(R.shape[:-2] + (9,)))


def quat_mul(a, b):
    return L.rowop(L.ROP_QUAT_MUL, a, b.expand_as(a))


def quat_conjugate(q):
    return torch.cat([q[..., :1], -q[..., 1:]], dim=-1)


def rot6d_to_rotmat(d6):
    return L.rowop(L.ROP_ROT6D_TO_ROTMAT, d6).reshape(d6.shape[:-1] + (3, 3))


def rotmat_to_rot6d(R):
    return torch.cat([R[..., 0], R[..., 1]], dim=-1)


def angle_axis_to_rot6d(aa):
    return rotmat_to_rot6d(angle_axis_to_rotation_matrix(aa))


def quat_to_rot6d(q):
    return rotmat_to_rot6d(quaternion_to_rotation_matrix(q))


def rot6d_to_quat(d6):
    return rotation_matrix_to_quaternion(rot6d_to_rotmat(d6))


def safe_atan2(y, x):
    return L.rowop(L.ROP_SAFE_ATAN2, torch.stack([y, x], dim=-1))[..., 0]


def normalize(x, eps=1e-9):
    return x / x.norm(dim=-1).clamp(min=eps).unsqueeze(-1)


def quat_angle_diff(a, b, eps=1e-6):
    q = quat_mul(a, quat_conjugate(b))
    return torch.acos((2 * q[..., 0] ** 2 - 1).clamp(-1 + eps, 1 - eps))


def get_heading(q):
    return 2 * safe_atan2(q[..., 3], q[..., 0])


def get_heading_q(q):
    z = torch.zeros_like(q[..., 0])
    return normalize(torch.stack([q[..., 0], z, z, q[..., 3]], dim=-1))


def heading_to_vec(h):
    return torch.stack([torch.cos(h), torch.sin(h)], dim=-1)


def vec_to_heading(v):
    return safe_atan2(v[..., 1].contiguous(), v[..., 0].contiguous())


def heading_to_quat(h):
    z = torch.zeros_like(h)
    return angle_axis_to_quaternion(torch.stack([z, z, h], dim=-1))


def deheading_quat(q, heading_q=None):
    if heading_q is None:
        heading_q = get_heading_q(q)
    return quat_mul(quat_conjugate(heading_q), q)


def make_transform(rot, trans, rot_type=None):
    if rot_type == 'axis_angle':
        rot = angle_axis_to_rotation_matrix(rot)
    elif rot_type == '6d':
        rot = rot6d_to_rotmat(rot)
    M = torch.eye(4, device=trans.device, dtype=trans.dtype).repeat(rot.shape[:-2] + (1, 1))
    M[..., :3, :3] = rot
    M[..., :3, 3] = trans
    return M


def transform_trans(M, x):
    while M.dim() < x.dim() + 1:
        M = M.unsqueeze(-3)
    xh = torch.cat([x, torch.ones_like(x[..., :1])], dim=-1).unsqueeze(-2)
    return torch.matmul(xh, M.transpose(-2, -1))[..., 0, :3]


def transform_rot(M, aa):
    R = angle_axis_to_rotation_matrix(aa)
    while M.dim() < R.dim():
        M = M.unsqueeze(-3)
    return rotation_matrix_to_angle_axis(torch.matmul(M[..., :3, :3], R).contiguous())


def inverse_transform(M):
    inv = torch.zeros_like(M)
    inv[..., :3, :3] = M[..., :3, :3].transpose(-2, -1)
    inv[..., :3, 3] = -torch.matmul(M[..., :3, 3].unsqueeze(-2), M[..., :3, :3]).squeeze(-2)
    inv[..., 3, 3] = 1.0
    return inv


def to34(M):
    """[...,4,4] -> contiguous [...,12] (3x4 row-major) as the CUDA library stores rigid transforms"""
    return M[..., :3, :].reshape(M.shape[:-2] + (12,)).contiguous()


def from34(m):
    M = torch.zeros(m.shape[:-1] + (4, 4), device=m.device, dtype=m.dtype)
    M[..., :3, :] = m.reshape(m.shape[:-1] + (3, 4))
    M[..., 3, 3] = 1.0
    return M
