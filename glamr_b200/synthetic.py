"""Seeded synthetic inputs for the global-reconstruction path.

There is no network in the build/bench environment and the SMPL body model is licence-gated, so every test
and benchmark runs on assets generated here.  Three generators:

* ``make_smpl_assets``   -- an SMPL-shaped body model (6890 vertices, 24 joints, 10 betas, 207 pose features)
* ``make_pose_dict``     -- one person's HybrIK-shaped ``pose_dict`` (layout written by the reference's
                            ``pose_est/hybrik_demo/demo.py:348-354`` and consumed at
                            ``global_recon/models/global_recon_model.py:88-125``)
* ``make_in_dict``       -- the ``in_dict`` handed to ``GlobalReconOptimizer.optimize`` (``run_demo.py:78-79``)

Everything is numpy on the host: this is data generation, not the product's compute path.
"""
import numpy as np
from scipy.spatial.transform import Rotation

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_FEAT = 207

# standard SMPL kinematic tree (kintree_table[0] of the model file)
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
                        dtype=np.int64)

# smplx VertexJointSelector picks appended after the 24 LBS joints (nose, eyes, ears, feet, then fingertips)
EXTRA_VERTEX_IDS = np.array([332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                             2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133], dtype=np.int64)

# index into [24 LBS | 21 picks | 9 extra-regressed] for the 26 'body26fk' joints
# (reference lib/models/smpl.py:35-57 JOINT_MAP composed with :221-250)
BODY26FK_JOINT_MAP = np.array([49, 1, 2, 51, 4, 5, 12, 7, 8, 29, 32, 30, 33, 31, 34, 24, 26, 25, 28, 27,
                               16, 17, 18, 19, 20, 21], dtype=np.int64)

# (body26fk index, SMPL joint index) pairs that share a joint name (global_recon_model.py:82-85)
SMPL_TO_BODY26FK = np.array([[0, 0], [1, 1], [2, 2], [4, 4], [5, 5], [6, 12], [7, 7], [8, 8],
                             [20, 16], [21, 17], [22, 18], [23, 19], [24, 20], [25, 21]], dtype=np.int64)


def make_smpl_assets(seed=0, skin_nnz=4, reg_nnz=8):
    """SMPL-shaped constants with the sparsity pattern of the real model (<=4 skin weights per vertex,
    a handful of vertices per regressor row), stored dense like the reference stores them."""
    rng = np.random.default_rng(seed)
    a = {}
    a['v_template'] = rng.normal(0.0, 0.3, (NUM_VERTS, 3)).astype(np.float32)
    a['shapedirs'] = rng.normal(0.0, 0.01, (NUM_VERTS, 3, NUM_BETAS)).astype(np.float32)
    a['posedirs'] = rng.normal(0.0, 0.001, (NUM_POSE_FEAT, NUM_VERTS * 3)).astype(np.float32)
    w = np.zeros((NUM_VERTS, NUM_JOINTS), np.float32)
    for v in range(NUM_VERTS):
        js = rng.choice(NUM_JOINTS, skin_nnz, replace=False)
        w[v, js] = rng.dirichlet(np.ones(skin_nnz)).astype(np.float32)
    a['lbs_weights'] = w

    def regressor(rows):
        r = np.zeros((rows, NUM_VERTS), np.float32)
        for i in range(rows):
            vs = rng.choice(NUM_VERTS, reg_nnz, replace=False)
            r[i, vs] = rng.dirichlet(np.ones(reg_nnz)).astype(np.float32)
        return r
    a['J_regressor'] = regressor(NUM_JOINTS)
    a['J_regressor_extra'] = regressor(9)
    a['parents'] = SMPL_PARENTS.copy()
    a['faces'] = rng.integers(0, NUM_VERTS, (13776, 3)).astype(np.int64)
    return a


def make_h36m_regressor(seed=0, nnz=6):
    """[17, 6890] joint regressor with the sparsity of the real J_regressor_h36m.npy (4-9 non-zeros per row, rows sum to 1)"""
    rng = np.random.default_rng(seed + 4242)
    r = np.zeros((17, NUM_VERTS), np.float32)
    for i in range(17):
        vs = rng.choice(NUM_VERTS, nnz, replace=False)
        r[i, vs] = rng.dirichlet(np.ones(nnz)).astype(np.float32)
    return r


def make_eval_case(num_persons=2, num_fr=60, seed=0, gaps=True):
    """A (result, ground truth) pair shaped like what Evaluator.compute_sequence_metrics receives (evaluator.py:218-327):
    `person_data[p]` with the optimiser's output keys, `gt[p]` with AMASS/3DPW-style `pose [T,72]`, `shape [10]`,
    `root_trans [T,3]`.  The estimate is the ground truth plus smooth errors; person 0 exists on a sub-range."""
    rng = np.random.default_rng(seed + 99)
    data = {'person_data': {}, 'gt': {}, 'gt_meta': {}, 'seq_len': num_fr, 'seq_name': 'eval_case'}
    for p in range(num_persons):
        T = num_fr
        pose = rng.normal(0.0, 0.2, (1, 72)) + np.cumsum(rng.normal(0.0, 0.01, (T, 72)), axis=0)
        pose[:, :3] = np.array([1.2, 1.2, 1.2]) + np.cumsum(rng.normal(0.0, 0.01, (T, 3)), axis=0)
        shape = rng.normal(0.0, 0.5, 10)
        trans = np.array([0.8 * p, 0.3, 0.9]) + np.cumsum(rng.normal(0.0, 0.01, (T, 3)), axis=0)
        vis = make_exist_with_gaps(T, seed=seed * 7 + p, n_gaps=2, min_len=4, max_len=10) if gaps else np.ones(T)
        exist = np.ones(T, dtype=bool)
        if p == 0 and T > 20:
            exist[:3] = False
            vis[:3] = 0
        est_pose = pose + rng.normal(0.0, 0.02, pose.shape) + np.cumsum(rng.normal(0.0, 0.002, pose.shape), axis=0)
        data['gt'][p] = {'pose': pose.astype(np.float32), 'shape': shape.astype(np.float32), 'root_trans': trans.astype(np.float32)}
        data['person_data'][p] = {
            'smpl_orient_world': est_pose[:, :3].astype(np.float32), 'smpl_pose': est_pose[:, 3:].astype(np.float32),
            'smpl_beta': np.repeat((shape + rng.normal(0, 0.1, 10))[None], T, axis=0).astype(np.float32),
            'root_trans_world': (trans + rng.normal(0.0, 0.01, trans.shape) + np.array([0.05, -0.02, 0.01])).astype(np.float32),
            'scale': None, 'visible_orig': vis.astype(np.float64), 'exist_frames': exist,
        }
    return data


def _rodrigues_np(aa):
    return Rotation.from_rotvec(aa.reshape(-1, 3)).as_matrix().reshape(aa.shape[:-1] + (3, 3))


def fk_joints_np(assets, pose_aa, betas):
    """24 posed LBS joints (rooted at the SMPL pelvis) for data generation only."""
    T = pose_aa.shape[0]
    v_shaped = assets['v_template'][None] + np.einsum('bl,mkl->bmk', betas, assets['shapedirs'])
    J = np.einsum('bik,ji->bjk', v_shaped, assets['J_regressor'])
    R = _rodrigues_np(pose_aa.reshape(T, 24, 3))
    parents = assets['parents']
    G_R = np.zeros((T, 24, 3, 3))
    G_t = np.zeros((T, 24, 3))
    G_R[:, 0] = R[:, 0]
    G_t[:, 0] = J[:, 0]
    for k in range(1, 24):
        p = parents[k]
        G_R[:, k] = G_R[:, p] @ R[:, k]
        G_t[:, k] = np.einsum('bij,bj->bi', G_R[:, p], J[:, k] - J[:, p]) + G_t[:, p]
    return G_t - G_t[:, [0]]


def make_pose_dict(assets, person, num_fr, seed=0, exist=None, kp_noise_px=2.0):
    """HybrIK-shaped estimates for one person.  Arrays hold VISIBLE frames only."""
    rng = np.random.default_rng(1000 * (seed + 1) + person)
    T = num_fr
    base = rng.normal(0.0, 0.2, (1, 24, 3))
    aa = base + np.cumsum(rng.normal(0.0, 0.01, (T, 24, 3)), axis=0)
    aa[:, 0] = np.array([np.pi, 0.0, 0.0]) + rng.normal(0.0, 0.05, (1, 3)) + np.cumsum(rng.normal(0.0, 0.005, (T, 3)), axis=0)
    betas = np.repeat(rng.normal(0.0, 0.5, (1, NUM_BETAS)), T, axis=0) + rng.normal(0.0, 0.01, (T, NUM_BETAS))
    root_trans = np.array([0.8 * person, 0.2, 5.0]) + np.cumsum(rng.normal(0.0, 0.01, (T, 3)), axis=0)
    cam_K = np.tile(np.array([[1000.0, 0.0, 960.0], [0.0, 1000.0, 540.0], [0.0, 0.0, 1.0]]), (T, 1, 1))

    joints = fk_joints_np(assets, aa.reshape(T, 72), betas) + root_trans[:, None]
    proj = np.einsum('bij,bkj->bki', cam_K, joints)
    kp24 = proj[..., :2] / proj[..., 2:]
    kp_2d = np.zeros((T, 29, 2))
    kp_2d[:, :24] = kp24 + rng.normal(0.0, kp_noise_px, kp24.shape)
    kp_2d[:, 24:] = kp24[:, :5]

    rotmats = _rodrigues_np(aa)                      # live reference code reads 24 rotation matrices
    if exist is None:
        exist = np.ones(T)
    exist = np.asarray(exist, dtype=np.float64)
    vis = exist == 1
    return {
        'smpl_pose_quat_wroot': rotmats.reshape(T, 54, 4).astype(np.float32)[vis],
        'smpl_beta': betas.astype(np.float32)[vis],
        'root_trans': root_trans.astype(np.float32)[vis],
        'kp_2d': kp_2d.astype(np.float32)[vis],
        'cam_K': cam_K.astype(np.float32)[vis],
        'frames': np.where(vis)[0],
        'frame2ind': {int(f): i for i, f in enumerate(np.where(vis)[0])},
        'bboxes_dict': {'id': person, 'exist': exist, 'start': int(np.where(vis)[0][0]),
                        'end': int(np.where(vis)[0][-1]), 'num_frames': int(vis.sum()),
                        'exist_frames': np.where(vis)[0]},
    }


def make_exist_with_gaps(num_fr, seed=0, n_gaps=2, min_len=10, max_len=60):
    """3DPW-like visibility: a few runs of invisible frames, never frame 0."""
    rng = np.random.default_rng(seed + 77)
    exist = np.ones(num_fr)
    for _ in range(n_gaps):
        ln = int(rng.integers(min_len, max_len + 1))
        ln = min(ln, max(1, num_fr // 4))
        s = int(rng.integers(11, max(12, num_fr - ln - 1)))
        exist[s:s + ln] = 0
    exist[0] = 1
    exist[-1] = 1
    return exist


def make_in_dict(assets, num_persons, num_fr, seed=0, gaps=False, seq_name='synthetic'):
    est = {}
    for p in range(num_persons):
        exist = make_exist_with_gaps(num_fr, seed=seed * 31 + p) if gaps else None
        est[p] = make_pose_dict(assets, p, num_fr, seed=seed, exist=exist)
    return {'est': est, 'gt': {}, 'gt_meta': {}, 'seq_name': seq_name}


class SyntheticPrior:
    """Stand-in for the learned motion/trajectory prior with the ``MotionTrajJointModel.inference`` contract
    (motion_infiller/models/motion_traj_joint_model.py:141-145): returns the input body pose as the "infilled" pose
    and a smooth seeded 11-D local trajectory.  Used only to seed benchmark problems when no network weights exist;
    both the CUDA arm and the CPU reference arm of bench.py consume the same instance's outputs."""

    def __init__(self, seed=0, device='cpu'):
        self.seed, self.device, self.calls = seed, device, 0

    def inference(self, batch, sample_num=1):
        import torch
        pose = batch['in_body_pose']
        B, T = pose.shape[0], pose.shape[1]
        rng = np.random.default_rng(self.seed + 17 * self.calls)
        self.calls += 1
        local = np.zeros((T, B, 11), np.float32)
        local[..., 0:2] = rng.normal(0.0, 0.01, (T, B, 2))
        local[..., 2] = 0.9 + np.cumsum(rng.normal(0.0, 0.002, (T, B)), axis=0)
        d6 = np.array([1.0, 0.0, 0.0, 0.0, 1.0, 0.0]) + np.cumsum(rng.normal(0.0, 0.01, (T, B, 6)), axis=0)
        local[..., 3:9] = d6
        dh = rng.normal(0.0, 0.02, (T, B))
        local[..., 9], local[..., 10] = np.cos(dh), np.sin(dh)
        t = lambda a: torch.tensor(a, dtype=torch.float32, device=self.device)
        return {'infer_out_body_pose': pose.to(torch.float32).reshape(B, 1, T, 69).clone(),
                'infer_out_local_traj_tp': t(local).reshape(T, B, 1, 11),
                'infer_out_orient': torch.zeros((B, 1, T, 3), dtype=torch.float32, device=self.device),
                'infer_out_trans': torch.zeros((B, 1, T, 3), dtype=torch.float32, device=self.device)}


class LatentInjector:
    """Wraps a MotionTrajJointModel-like object and injects seeded latents through the reference's own injection
    points (`in_motion_latent`, `in_traj_latent`) so CPU and CUDA runs sample the same z."""

    def __init__(self, model, seed=0):
        self.model, self.seed, self.calls = model, seed, 0

    @property
    def supports_person_batch(self):
        return getattr(self.model, 'supports_person_batch', False)

    def inference(self, batch, sample_num=1):
        import torch
        B, T = batch['in_body_pose'].shape[:2]
        dev = batch['in_body_pose'].device
        zm, zt = [], []
        for _ in range(B):                       # sequence b of a batched call draws what the b-th single call would have
            g = torch.Generator().manual_seed(self.seed + 101 * self.calls)
            self.calls += 1
            zm.append(torch.randn(int(np.ceil((T - 10) / 30)), 128, generator=g))
            zt.append(torch.randn(1, 128, generator=g))
        b = dict(batch)
        if B == 1:
            b['in_motion_latent'], b['in_traj_latent'] = zm[0].to(dev), zt[0].to(dev)
        else:
            b['in_motion_latent'], b['in_traj_latent'] = torch.stack(zm).to(dev), torch.cat(zt).to(dev)
        return self.model.inference(b, sample_num=sample_num)
