"""Import shim: exposes the smplx.lbs functions GLAMR imports, taken from the smplx-derived copy that is
vendored inside the reference tree (HybrIK/hybrik/models/layers/smpl/lbs.py).  Used ONLY by the golden-vector
generator in this container; it never travels to the GPU box and no product code imports it."""
import importlib.util
import os
import torch

_REF = os.environ.get('GLAMR_REFERENCE_ROOT', '/root/reference')
_spec = importlib.util.spec_from_file_location('_hybrik_lbs', f'{_REF}/HybrIK/hybrik/models/layers/smpl/lbs.py')
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
vertices2joints, blend_shapes = _mod.vertices2joints, _mod.blend_shapes
batch_rodrigues, batch_rigid_transform, transform_mat = _mod.batch_rodrigues, _mod.batch_rigid_transform, _mod.transform_mat


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True):
    B = max(betas.shape[0], pose.shape[0])
    dev, dt = betas.device, betas.dtype
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    rot_mats = batch_rodrigues(pose.view(-1, 3)).view([B, -1, 3, 3])
    pose_feature = (rot_mats[:, 1:] - torch.eye(3, dtype=dt, device=dev)).view([B, -1])
    v_posed = torch.matmul(pose_feature, posedirs).view(B, -1, 3) + v_shaped
    J_t, A = batch_rigid_transform(rot_mats, J, parents, dtype=dt)
    T = torch.matmul(lbs_weights.unsqueeze(0).expand([B, -1, -1]), A.view(B, J_regressor.shape[0], 16)).view(B, -1, 4, 4)
    v_h = torch.cat([v_posed, torch.ones([B, v_posed.shape[1], 1], dtype=dt, device=dev)], dim=2)
    return torch.matmul(T, v_h.unsqueeze(-1))[:, :, :3, 0], J_t
