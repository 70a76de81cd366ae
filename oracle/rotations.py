"""ORACLE (test infrastructure, not product code): CPU restatement in plain torch of the rotation / rigid-transform
algebra the GLAMR global-reconstruction path uses.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this package.

Each function states the reference location it follows (paths relative to /root/reference).  Quaternions are
WXYZ.  The exact eps / clamp / branch conventions matter because the CUDA kernels reproduce both the values and
autograd's derivative of these very formulas.  Pinned against the executed
reference by tests/golden (see tests/golden/make_golden.py).
"""
import torch


# ----------------------------------------------------------------------------- small helpers
def unit(x, eps=1e-9):
    """lib/utils/torch_transform.py:6-7  x / max(|x|, eps)"""
    return x / x.norm(dim=-1).clamp(min=eps).unsqueeze(-1)


def safe_atan2(y, x, eps=1e-6):
    """lib/utils/torch_transform.py:63-67 (same in konia_transform.py:44-47): nudge y by eps where both
    arguments are tiny; the nudge is a constant so the derivative is that of atan2 at the nudged point."""
    nudge = ((y.abs() < eps) & (x.abs() < eps)).to(y.dtype) * eps
    return torch.atan2(y + nudge, x)


def _safe_div(num, den, eps=1e-6):
    """lib/utils/konia_transform.py:343-346"""
    return num / (den + (den.abs() < eps).to(den.dtype) * eps)


# ----------------------------------------------------------------------------- quaternion algebra
def quat_mul(a, b):
    """lib/utils/torch_transform.py:10-28 -- the 8-multiplication Hamilton product, kept in the same operation
    order so fp32 rounding matches."""
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    return torch.stack([qq - ww + (z1 - y1) * (y2 - z2),
                        qq - xx + (x1 + w1) * (x2 + w2),
                        qq - yy + (w1 - x1) * (y2 + z2),
                        qq - zz + (z1 + y1) * (w2 - x2)], dim=-1)


def quat_conj(q):
    """lib/utils/torch_transform.py:31-35"""
    return torch.cat([q[..., :1], -q[..., 1:]], dim=-1)


def quat_angle(q, eps=1e-6):
    """lib/utils/torch_transform.py:48-55"""
    return torch.acos((2 * q[..., 0] ** 2 - 1).clamp(-1 + eps, 1 - eps))


def quat_angle_diff(a, b):
    """lib/utils/torch_transform.py:58-60"""
    return quat_angle(quat_mul(a, quat_conj(b)))


def get_heading(q, eps=1e-6):
    """lib/utils/torch_transform.py:172-177"""
    return 2 * safe_atan2(q[..., 3], q[..., 0], eps)


def get_heading_q(q):
    """lib/utils/torch_transform.py:180-185"""
    z = torch.zeros_like(q[..., 0])
    return unit(torch.stack([q[..., 0], z, z, q[..., 3]], dim=-1))


def heading_to_vec(h):
    """lib/utils/torch_transform.py:188-191"""
    return torch.stack([torch.cos(h), torch.sin(h)], dim=-1)


def vec_to_heading(v):
    """lib/utils/torch_transform.py:194-197"""
    return safe_atan2(v[..., 1], v[..., 0])


def heading_to_quat(h):
    """lib/utils/torch_transform.py:200-204"""
    z = torch.zeros_like(h)
    return aa_to_quat(torch.stack([z, z, h], dim=-1))


def deheading_quat(q, heading_q=None):
    """lib/utils/torch_transform.py:207-211"""
    if heading_q is None:
        heading_q = get_heading_q(q)
    return quat_mul(quat_conj(heading_q), q)


# ----------------------------------------------------------------------------- conversions (kornia-derived)
def aa_to_rotmat(aa):
    """lib/utils/konia_transform.py:234-313.  Both branches are evaluated and blended with a 0/1 mask, so the
    gradient is mask*d(normal) + (1-mask)*d(taylor)."""
    shp = aa.shape
    r = aa.reshape(-1, 3)
    th2 = (r * r).sum(-1, keepdim=True)
    th = torch.sqrt(th2.clamp_min(1e-6))
    w = r / (th + 1e-6)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = torch.cos(th), torch.sin(th)
    one_c = 1.0 - c
    normal = torch.cat([c + wx * wx * one_c, wx * wy * one_c - wz * s, wy * s + wx * wz * one_c,
                        wz * s + wx * wy * one_c, c + wy * wy * one_c, -wx * s + wy * wz * one_c,
                        -wy * s + wx * wz * one_c, wx * s + wy * wz * one_c, c + wz * wz * one_c], dim=1)
    rx, ry, rz = r[:, 0:1], r[:, 1:2], r[:, 2:3]
    one = torch.ones_like(rx)
    taylor = torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1)
    m = (th2 > 1e-6).to(r.dtype)
    return (m * normal + (1.0 - m) * taylor).reshape(shp[:-1] + (3, 3))


def rotmat_to_quat(R, eps=1e-6):
    """lib/utils/konia_transform.py:349-443: four candidates, nested torch.where selection, no normalisation."""
    m = R.reshape(R.shape[:-2] + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[..., i:i + 1] for i in range(9)]
    tr = m00 + m11 + m22

    def cand(diag, a, b, c, order):
        sq = torch.sqrt(diag.clamp_min(eps)) * 2.0
        comps = {'s': 0.25 * sq, 'a': _safe_div(a, sq), 'b': _safe_div(b, sq), 'c': _safe_div(c, sq)}
        return torch.cat([comps[k] for k in order], dim=-1)

    q_tr = cand(tr + 1.0, m21 - m12, m02 - m20, m10 - m01, 'sabc')
    q_x = cand(1.0 + m00 - m11 - m22, m21 - m12, m01 + m10, m02 + m20, 'asbc')
    q_y = cand(1.0 + m11 - m00 - m22, m02 - m20, m01 + m10, m12 + m21, 'absc')
    q_z = cand(1.0 + m22 - m00 - m11, m10 - m01, m02 + m20, m12 + m21, 'abcs')
    q = torch.where(m11 > m22, q_y, q_z)
    q = torch.where((m00 > m11) & (m00 > m22), q_x, q)
    return torch.where(tr > 0.0, q_tr, q)


def quat_to_aa(q, eps=1e-6):
    """lib/utils/konia_transform.py:560-630"""
    w, x, y, z = q.unbind(-1)
    s2 = x * x + y * y + z * z
    s = torch.sqrt(s2.clamp_min(eps))
    two_theta = 2.0 * torch.where(w < 0.0, safe_atan2(-s, -w), safe_atan2(s, w))
    k = torch.where(s2 > 0.0, _safe_div(two_theta, s, eps), 2.0 * torch.ones_like(s))
    return torch.stack([x * k, y * k, z * k], dim=-1)


def aa_to_quat(aa, eps=1e-6):
    """lib/utils/konia_transform.py:753-822"""
    th2 = (aa * aa).sum(-1, keepdim=True)
    th = torch.sqrt(th2.clamp_min(eps))
    half = th * 0.5
    pos = th2 > 0.0
    k = torch.where(pos, _safe_div(torch.sin(half), th, eps), 0.5 * torch.ones_like(th))
    w = torch.where(pos, torch.cos(half), torch.ones_like(th))
    return torch.cat([w, aa * k], dim=-1)


def quat_to_rotmat(q):
    """lib/utils/konia_transform.py:477-557 (normalises first, eps 1e-12)"""
    qn = torch.nn.functional.normalize(q, p=2.0, dim=-1, eps=1e-12)
    w, x, y, z = qn.unbind(-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = 1.0
    return torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                        txy + twz, one - (txx + tzz), tyz - twx,
                        txz - twy, tyz + twx, one - (txx + tyy)], dim=-1).reshape(q.shape[:-1] + (3, 3))


def rotmat_to_aa(R):
    """lib/utils/konia_transform.py:316-339"""
    return quat_to_aa(rotmat_to_quat(R))


def rotmat_to_rot6d(R):
    """lib/utils/torch_transform.py:214-217: first two COLUMNS"""
    return torch.cat([R[..., 0], R[..., 1]], dim=-1)


def rot6d_to_rotmat(d6):
    """lib/utils/torch_transform.py:220-227: Gram-Schmidt, columns (b1, b2, b1 x b2)"""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = unit(a1)
    b2 = unit(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack([b1, b2, b3], dim=-1)


def aa_to_rot6d(aa):
    """lib/utils/torch_transform.py:230-231"""
    return rotmat_to_rot6d(aa_to_rotmat(aa))


def quat_to_rot6d(q):
    """lib/utils/torch_transform.py:238-239"""
    return rotmat_to_rot6d(quat_to_rotmat(q))


def rot6d_to_quat(d6):
    """lib/utils/torch_transform.py:242-243"""
    return rotmat_to_quat(rot6d_to_rotmat(d6))


# ----------------------------------------------------------------------------- 4x4 rigid transforms
def make_transform(rot, trans, rot_type=None):
    """lib/utils/torch_transform.py:246-254"""
    if rot_type == 'axis_angle':
        rot = aa_to_rotmat(rot)
    elif rot_type == '6d':
        rot = rot6d_to_rotmat(rot)
    top = torch.cat([rot, trans.unsqueeze(-1)], dim=-1)
    bottom = torch.zeros_like(top[..., :1, :])
    bottom = bottom + torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=top.dtype, device=top.device)
    return torch.cat([top, bottom], dim=-2)


def transform_trans(M, x):
    """lib/utils/torch_transform.py:257-262: (M @ [x;1])[:3], M broadcast over extra point dims"""
    while M.dim() < x.dim() + 1:
        M = M.unsqueeze(-3)
    xh = torch.cat([x, torch.ones_like(x[..., :1])], dim=-1).unsqueeze(-2)
    return torch.matmul(xh, M.transpose(-2, -1))[..., 0, :3]


def transform_rot(M, aa):
    """lib/utils/torch_transform.py:265-271"""
    R = aa_to_rotmat(aa)
    while M.dim() < R.dim():
        M = M.unsqueeze(-3)
    return rotmat_to_aa(torch.matmul(M[..., :3, :3], R))


def inverse_transform(M):
    """lib/utils/torch_transform.py:274-279"""
    Rt = M[..., :3, :3].transpose(-2, -1)
    t = -torch.matmul(M[..., :3, 3].unsqueeze(-2), M[..., :3, :3]).squeeze(-2)
    top = torch.cat([Rt, t.unsqueeze(-1)], dim=-1)
    bottom = torch.zeros_like(top[..., :1, :])
    bottom = bottom + torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=top.dtype, device=top.device)
    return torch.cat([top, bottom], dim=-2)


def perspective_projection(p3d, K):
    """lib/utils/geometry.py:23-25"""
    p = torch.matmul(K, p3d.transpose(2, 1)).transpose(2, 1)
    return p[:, :, :2] / (p[:, :, 2:] + 1e-8)
