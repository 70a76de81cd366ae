"""ORACLE (test infrastructure): the residual terms of GLAMR's global optimisation, restated from
global_recon/models/loss_func.py.  Every term maps (data, specs) -> 0-d tensor, summed over persons with the
reference's normalisers.  `kp_2d*` inputs are float64 in the reference
(global_recon_model.py:119), so that term is evaluated in float64 here as well.
"""
import torch

from . import rotations as rt

FPS = 30.0


def _persons(data):
    return data['person_data'].values()


def kp_2d(data, specs):
    """loss_func.py:6-36  Geman-McClure (sigma 100) reprojection error, score^2-weighted."""
    min_conf = specs.get('min_conf', 0.05)
    first_only = specs.get('first_frame_only', False)
    ffw = specs.get('first_frame_weight', 1.0)
    total, n = 0, 0
    for pd in _persons(data):
        vis = pd['vis_frames']
        d = pd['kp_2d_pred'][vis] - pd['kp_2d_aligned'][vis]
        conf = pd['kp_2d_score'][vis].clone()
        conf = torch.where(conf < min_conf, torch.zeros_like(conf), conf)
        sq = d * d
        rho = (100.0 ** 2 * sq) / (100.0 ** 2 + sq)
        if first_only:
            rho = rho[:1]
        n = n + vis.sum()
        w = torch.ones(rho.shape[0], dtype=rho.dtype)
        w[:10] = ffw
        rho = rho * w[:, None, None]
        # with first_frame_only rho is [1,26,2] and broadcasts against ALL visible frames' scores (reference quirk)
        total = total + (rho.sum(-1) * conf ** 2).sum()
    return total / n


def kp_2d_dist(data, specs):
    """loss_func.py:39-57 (monitor): mean pixel distance over joints whose INPUT score exceeds min_conf."""
    min_conf = specs.get('min_conf', 0.05)
    first_only = specs.get('first_frame_only', False)
    dists = []
    for pd in _persons(data):
        sc, pr, al = pd['kp_2d_score'], pd['kp_2d_pred'], pd['kp_2d_aligned']
        if first_only:
            sc, pr, al = sc[:1], pr[:1], al[:1]
        keep = (sc > min_conf).reshape(-1)
        dists.append((pr - al).pow(2).sum(-1).sqrt().reshape(-1)[keep])
    return torch.cat(dists).mean()


def cam_rot_smoothness(data, specs):
    """loss_func.py:60-65"""
    v = (data['cam_rot_6d'][:-1] - data['cam_rot_6d'][1:]) * FPS
    return v.pow(2).sum(-1).mean()


def cam_trans_smoothness(data, specs):
    """loss_func.py:68-73"""
    v = (data['cam_trans'][:-1] - data['cam_trans'][1:]) * FPS
    return v.pow(2).sum(-1).mean()


def cam_inv_rot_smoothness(data, specs):
    """loss_func.py:76-81"""
    c = data['cam_pose_inv'][:, :3, :2]
    return ((c[:-1] - c[1:]) * FPS).pow(2).sum(-1).sum(-1).mean()


def cam_origin_smoothness(data, specs):
    """loss_func.py:84-91"""
    o = data['cam_pose_inv'][:, :3, 3]
    return ((o[1:] - o[:-1]) * FPS).pow(2).sum(-1).mean()


def cam_depth_smoothness(data, specs):
    """loss_func.py:94-103"""
    c = data['cam_pose_inv']
    dz = ((c[:-1, :3, 3] - c[1:, :3, 3]) * c[1:, :3, 2]).sum(-1) * FPS
    return dz.pow(2).sum(-1).mean()


def cam_up_reg(data, specs):
    """loss_func.py:106-114 -- linear in cam_pose_inv[:, 2, 1]"""
    ffw = specs.get('first_frame_weight', 1.0)
    u = data['cam_pose_inv'][:, 2, 1]
    w = torch.ones_like(u)
    w[:10] = ffw
    u = u * w
    if specs.get('first_frame_only', False):
        u = u[:1]
    return u.mean()


def traj_rot_smoothness(data, specs):
    """loss_func.py:117-132"""
    rot_type = specs.get('rot_type', '6d')
    total, n = 0, 0
    for pd in _persons(data):
        o = pd['smpl_orient_world']
        n += o.shape[0] - 1
        if rot_type == '6d':
            r = rt.aa_to_rot6d(o)
            diff = r[1:] - r[:-1]
        else:
            q = rt.aa_to_quat(o)
            diff = rt.quat_angle_diff(q[1:], q[:-1])
        total = total + (diff * FPS).pow(2).sum()
    return total / n


def traj_trans_smoothness(data, specs):
    """loss_func.py:135-144"""
    total, n = 0, 0
    for pd in _persons(data):
        t = pd['root_trans_world']
        n += t.shape[0] - 1
        total = total + ((t[1:] - t[:-1]) * FPS).pow(2).sum()
    return total / n


def cam_traj_rot(data, specs):
    """loss_func.py:147-172"""
    rot_type = specs.get('rot_type', '6d')
    ffw = specs.get('first_frame_weight', 1.0)
    first_only = specs.get('first_frame_only', False)
    total, n = 0, 0
    for pd in _persons(data):
        vis = pd['vis_frames']
        a, b = pd['smpl_orient_cam_in_world'][vis], pd['smpl_orient_cam'][vis]
        if rot_type == '6d':
            diff = rt.aa_to_rot6d(b) - rt.aa_to_rot6d(a)
        else:
            diff = rt.quat_angle_diff(rt.aa_to_quat(b), rt.aa_to_quat(a))
        if first_only:
            diff = diff[:1]
            n = n + 1
        else:
            w = torch.ones(diff.shape[0], dtype=diff.dtype)
            w[0] = ffw
            diff = diff * (w[:, None] if diff.dim() == 2 else w)
            n = n + vis.sum()
        total = total + diff.pow(2).sum()
    return total / n


def cam_traj_trans(data, specs):
    """loss_func.py:175-186"""
    ffw = specs.get('first_frame_weight', 1.0)
    total, n = 0, 0
    for pd in _persons(data):
        vis = pd['vis_frames']
        n = n + vis.sum()
        diff = pd['root_trans_cam_in_world'][vis] - pd['root_trans_cam'][vis]
        w = torch.ones(diff.shape[0], dtype=diff.dtype)
        w[0] = ffw
        total = total + (diff * w[:, None]).pow(2).sum()
    return total / n


def _reg_person(key):
    """loss_func.py:189-196"""
    def fn(data, specs):
        total, n = 0, 0
        for pd in _persons(data):
            n += pd[key].shape[0]
            total = total + (pd[key] * FPS).pow(2).sum()
        return total / n
    return fn


def _reg_global(key):
    """loss_func.py:199-201"""
    def fn(data, specs):
        return (data[key] * FPS).pow(2).sum() / data[key].shape[0]
    return fn


def local_traj_dheading_reg_new(data, specs):
    """loss_func.py:220-230"""
    total, n = 0, 0
    for pd in _persons(data):
        dh = pd['traj_local_dheading']
        n += dh.shape[0]
        diff = rt.heading_to_vec(dh) - torch.tensor([1.0, 0.0], dtype=dh.dtype)
        total = total + (diff * FPS).pow(2).sum()
    return total / n


def rel_transform(data, specs):
    """loss_func.py:248-271 -- note it reads `first_frame_trans_only`, not `first_frame_only`."""
    tw = specs.get('trans_weight', 1.0)
    ffw = specs.get('first_frame_weight', 10)
    trans_first_only = specs.get('first_frame_trans_only', False)
    persons = data['person_data']
    total, n = 0, 0
    for (i, j), C in data['rel_transform_cam'].items():
        n += C.shape[0]
        both = persons[i]['vis_frames'] & persons[j]['vis_frames']
        if both.sum() == 0:
            continue
        W = torch.matmul(rt.inverse_transform(persons[i]['person_transform_world'][both]),
                         persons[j]['person_transform_world'][both])
        Cv = C[both]
        w = torch.ones(W.shape[0], dtype=W.dtype)
        w[0] = ffw
        d_rot = (Cv[..., :3, :2] - W[..., :3, :2]) * w[:, None, None]
        d_tr = (Cv[..., :3, 3] - W[..., :3, 3]) * w[:, None]
        if trans_first_only:
            keep = torch.zeros_like(w)
            keep[0] = 1.0
            d_tr = d_tr * keep[:, None]
        total = total + d_rot.pow(2).sum() + d_tr.pow(2).sum() * tw
    return total / n if n > 0 else total


def _latent_reg(key):
    """loss_func.py:293-310"""
    def fn(data, specs):
        total, n = 0, 0
        for pd in _persons(data):
            n += pd[key].shape[0]
            total = total + pd[key].pow(2).sum()
        return total / n
    return fn


# same registry keys as loss_func.py:314-340 ('penetration' needs the external `sdf` package: out of scope)
RESIDUALS = {
    'kp_2d': kp_2d,
    'kp_2d_dist': kp_2d_dist,
    'cam_rot_smoothness': cam_rot_smoothness,
    'cam_trans_smoothness': cam_trans_smoothness,
    'cam_inv_rot_smoothness': cam_inv_rot_smoothness,
    'cam_origin_smoothness': cam_origin_smoothness,
    'cam_depth_smoothness': cam_depth_smoothness,
    'traj_rot_smoothness': traj_rot_smoothness,
    'traj_trans_smoothness': traj_trans_smoothness,
    'cam_up_reg': cam_up_reg,
    'cam_traj_rot': cam_traj_rot,
    'cam_traj_trans': cam_traj_trans,
    'traj_rot_res': _reg_person('smpl_orient_world_res'),
    'traj_trans_res': _reg_person('root_trans_world_res'),
    'local_traj_dxy_reg': _reg_person('traj_local_dxy'),
    'local_traj_dheading_reg': _reg_person('traj_local_dheading'),
    'local_traj_dheading_reg_new': local_traj_dheading_reg_new,
    'local_traj_rot_reg': _reg_person('traj_local_rot'),
    'local_traj_z_reg': _reg_person('traj_local_z'),
    'cam_inv_trans_residual_reg': _reg_global('cam_inv_trans_residual'),
    'person2cam_res_trans_reg': _reg_global('person2cam_res_trans'),
    'rel_transform': rel_transform,
    'motion_latent_reg': _latent_reg('motion_latent'),
    'traj_latent_reg': _latent_reg('traj_latent'),
}
