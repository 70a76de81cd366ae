"""ORACLE (test infrastructure): torch-CPU restatement of the SMPL body model as GLAMR evaluates it.

The arithmetic lives in the third-party package `smplx` (unpinned: reference requirements.txt:1), which is absent
from /root/reference; its published algorithm is restated here from the smplx-derived copy vendored in the
reference tree, HybrIK/hybrik/models/layers/smpl/lbs.py:195-288,402-548, and from GLAMR's wrapper
lib/models/smpl.py:274-343.  The 21 `VertexJointSelector` vertex ids are recalled from upstream smplx and cannot
be verified offline; only 11 of them reach the body26fk joints and all of those carry zero
confidence in the loss.
"""
import numpy as np
import torch

from glamr_b200.synthetic import EXTRA_VERTEX_IDS, BODY26FK_JOINT_MAP


def rodrigues_smplx(rv):
    """HybrIK/.../lbs.py:446-477: angle = |r + 1e-8| (eps added per component), no small-angle branch."""
    angle = torch.norm(rv + 1e-8, dim=1, keepdim=True)
    d = rv / angle
    c, s = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    o = torch.zeros_like(x)
    K = torch.cat([o, -z, y, z, o, -x, -y, x, o], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rv.dtype, device=rv.device)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def rigid_chain(R, J, parents):
    """HybrIK/.../lbs.py:493-548.  R [B,24,3,3], J [B,24,3] -> posed joints [B,24,3], A [B,24,4,4]."""
    B, n = J.shape[:2]
    rel = J.clone()
    rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
    M = torch.zeros(B, n, 4, 4, dtype=J.dtype, device=J.device)
    M[..., :3, :3] = R
    M[..., :3, 3] = rel
    M[..., 3, 3] = 1.0
    G = [M[:, 0]]
    for k in range(1, n):
        G.append(torch.matmul(G[int(parents[k])], M[:, k]))
    G = torch.stack(G, dim=1)
    posed = G[..., :3, 3]
    Jh = torch.cat([J, torch.zeros_like(J[..., :1])], dim=-1).unsqueeze(-1)
    corr = torch.matmul(G, Jh)                               # [B,n,4,1]
    A = G - torch.cat([torch.zeros_like(G[..., :3]), corr], dim=-1)
    return posed, A


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """HybrIK/.../lbs.py:195-288 (without the h36m regressor): vertices [B,V,3], posed LBS joints [B,24,3]."""
    B = pose.shape[0]
    v_shaped = v_template[None] + torch.einsum('bl,mkl->bmk', betas, shapedirs)
    J = torch.einsum('bik,ji->bjk', v_shaped, J_regressor)
    R = rodrigues_smplx(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    feat = (R[:, 1:] - torch.eye(3, dtype=pose.dtype, device=pose.device)).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(feat, posedirs).view(B, -1, 3)
    posed, A = rigid_chain(R, J, parents)
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.reshape(B, -1, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones_like(v_posed[..., :1])], dim=-1)
    verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, posed


class OracleSMPL:
    """lib/models/smpl.py:274-343 on top of smplx.SMPL.forward (create_transl=False)."""

    def __init__(self, assets, device='cpu', dtype=torch.float32):
        t = lambda k: torch.tensor(np.asarray(assets[k]), dtype=dtype, device=device)
        self.v_template, self.shapedirs, self.posedirs = t('v_template'), t('shapedirs'), t('posedirs')
        self.J_regressor, self.lbs_weights, self.J_regressor_extra = t('J_regressor'), t('lbs_weights'), t('J_regressor_extra')
        self.parents = torch.tensor(np.asarray(assets['parents']), dtype=torch.long)
        self.joint_map = torch.tensor(BODY26FK_JOINT_MAP, dtype=torch.long)
        self.extra_v = torch.tensor(EXTRA_VERTEX_IDS, dtype=torch.long)

    def __call__(self, global_orient, body_pose, betas, root_trans=None, root_scale=None, orig_joints=False):
        if global_orient is None:
            global_orient = torch.zeros_like(body_pose[:, :3])
        pose = torch.cat([global_orient, body_pose], dim=1)
        verts, posed = lbs(betas, pose, self.v_template, self.shapedirs, self.posedirs, self.J_regressor,
                           self.parents, self.lbs_weights)
        j45 = torch.cat([posed, verts[:, self.extra_v]], dim=1)
        if orig_joints:
            joints = j45[:, :24]
        else:
            extra = torch.einsum('bik,ji->bjk', verts, self.J_regressor_extra)
            joints = torch.cat([j45, extra], dim=1)[:, self.joint_map]
        if root_trans is not None:
            scale = torch.ones_like(root_trans[:, 0]) if root_scale is None else root_scale
            root = joints[:, [0]]
            verts = (verts - root) * scale[:, None, None] + root_trans[:, None]
            joints = (joints - root) * scale[:, None, None] + root_trans[:, None]
        return joints, verts

    def get_joints(self, global_orient, body_pose, root_trans=None, root_scale=None):
        """lib/models/smpl.py:318-343: FK only, rest joints from v_template (betas ignored)."""
        pose = torch.cat([global_orient, body_pose], dim=1)
        B = pose.shape[0]
        J = torch.matmul(self.J_regressor, self.v_template)[None].repeat(B, 1, 1)
        R = rodrigues_smplx(pose.reshape(-1, 3)).view(B, -1, 3, 3)
        joints, _ = rigid_chain(R, J, self.parents)
        if root_trans is not None:
            scale = torch.ones_like(root_trans[:, 0]) if root_scale is None else root_scale
            joints = (joints - joints[:, [0]]) * scale[:, None, None] + root_trans[:, None]
        return joints
