"""ORACLE (test infrastructure): torch-CPU restatement of GLAMR's GlobalReconOptimizer
(global_recon/models/global_recon_model.py), i.e. init_data -> per-iteration forward -> residuals -> autograd
backward -> torch.optim.Adam.  It is the checker for the CUDA path and the "port" CPU baseline of bench.py; the
product (glamr_b200/) never imports it.

Pinned against the executed reference: tests/golden/make_golden.py runs the unmodified reference through the
import shims of oracle/refshim and stores its outputs; tests/test_oracle_vs_golden.py replays them here.

Not restated (raise NotImplementedError): latent optimisation (flag_opt_motion_latent / flag_opt_traj_latent),
the external-SDF penetration term, heading_type='vec', absolute_heading, flag_traj_from_cam -- none is enabled in
any shipped config.
"""
import time
import numpy as np
import torch
from scipy.interpolate import interp1d
from scipy.spatial.transform import Rotation

from glamr_b200.synthetic import SMPL_TO_BODY26FK
from . import rotations as rt
from . import traj_codec as tc
from .residuals import RESIDUALS
from .smpl import OracleSMPL


def _to_torch(x):
    if isinstance(x, np.ndarray):
        return torch.tensor(x)
    if isinstance(x, (np.floating, np.integer, np.bool_)):
        return torch.tensor(x)
    if isinstance(x, dict):
        return {k: _to_torch(v) for k, v in x.items()}
    return x


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy(v) for v in x)
    return x


def cast_floats(x, dtype):
    """every floating tensor of a nested data dict -> `dtype` (bool / integer tensors and python values untouched)"""
    if isinstance(x, torch.Tensor):
        return x.detach().to(dtype) if x.is_floating_point() else x
    if isinstance(x, dict):
        return {k: cast_floats(v, dtype) for k, v in x.items()}
    return x


class OracleGlobalRecon:
    """Same constructor/methods as the reference class (global_recon_model.py:23-67,572-589)."""

    def __init__(self, cfg, smpl_assets, mt_model=None, log=None):
        self.cfg = cfg
        self.specs = s = cfg.grecon_model_specs
        self.log = log
        self.cur_iter = 0
        self._assets = smpl_assets
        self.smpl = OracleSMPL(smpl_assets)
        self.mt_model = mt_model
        g = s.get
        self.est_type = g('est_type', 'hybrik')
        self.flag_infer_motion_traj = g('flag_infer_motion_traj', False)
        self.flag_infill_motion = g('flag_infill_motion', True)
        self.flag_pred_traj = g('flag_pred_traj', True)
        self.flag_opt_traj = g('flag_opt_traj', True)
        self.flag_opt_cam = g('flag_opt_cam', True)
        self.flag_fixed_cam = g('flag_fixed_cam', False)
        self.flag_opt_vis_local_rot = g('flag_opt_vis_local_rot', False)
        self.flag_opt_person2cam_rot = g('flag_opt_person2cam_rot', False)
        self.flag_opt_person2cam_trans = g('flag_opt_person2cam_trans', False)
        self.flag_cam_inv_trans_res_all = g('flag_cam_inv_trans_res_all', True)
        self.flag_filter_pose = g('flag_filter_pose', True)
        self.flag_make_invis_with_keypoint = g('flag_make_invis_with_keypoint', False)
        self.make_invis_keypoint_min_score = g('make_invis_keypoint_min_score', 0.6)
        self.make_invis_keypoint_min_num = g('make_invis_keypoint_min_num', 15)
        self.flag_opt_cam_from_person_pose = g('flag_opt_cam_from_person_pose', False)
        self.flag_init_cam_all_frames = g('flag_init_cam_all_frames', False)
        self.cam_fix_frames = g('cam_fix_frames', [[0, None]])
        self.opt_stage_specs = cfg.opt_stage_specs
        for flag in ['flag_opt_motion_latent', 'flag_opt_traj_latent', 'flag_use_pen_loss', 'flag_traj_from_cam',
                     'absolute_heading']:
            if g(flag, False):
                raise NotImplementedError(f'{flag} is not restated in the oracle')
        if g('heading_type', 'scalar') != 'scalar':
            raise NotImplementedError('heading_type != scalar')

    def to_float64(self, data):
        """Noise-floor instrument (tests/golden/make_golden.py): continue from the SAME float32 state with every
        per-iteration formula evaluated in float64 -- forward, residuals, autograd, Adam.  |ref32 - ref64| is then the
        rounding-noise amplification of the reference itself, the yardstick for |cuda - ref64|."""
        self.smpl = OracleSMPL(self._assets, dtype=torch.float64)
        return cast_floats(data, torch.float64)

    # ------------------------------------------------------------------ init (global_recon_model.py:76-248)
    def _person_from_estimate(self, est, gt_entry):
        d = {}
        visible = est['bboxes_dict']['exist'].copy()
        d['visible'] = visible
        d['visible_orig'] = visible.copy()
        where = np.where(visible)[0]
        start, end = where[0], where[-1] + 1
        d['fr_start'], d['fr_end'] = start, end
        exist = visible == 1
        exist[start:end] = True
        d['exist_frames'] = exist
        d['exist_len'] = end - start
        d['max_len'] = n = visible.shape[0]
        d['frames'] = np.arange(n)
        d['vis_frames'] = vis = visible == 1
        d['invis_frames'] = visible == 0
        d['frame2ind'] = {f: i for i, f in enumerate(d['frames'])}
        d['scale'] = None
        rotmats = est['smpl_pose_quat_wroot']
        nv = rotmats.shape[0]
        aa = Rotation.from_matrix(rotmats.reshape(-1, 3, 3)).as_rotvec().reshape(nv, -1, 3).astype(np.float32)
        d['smpl_pose'] = aa[:, 1:].reshape(-1, 69)
        if gt_entry is not None:
            d['smpl_pose_gt'] = gt_entry['pose'][:, 3:]
        d['smpl_beta'] = est['smpl_beta']
        d['smpl_orient_cam'] = aa[:, 0]
        d['root_trans_cam'] = est['root_trans']
        j2d = est['kp_2d'][:, :24]
        j2d = np.concatenate([j2d, np.ones_like(j2d[:, :, :1])], axis=-1)
        kp = np.zeros((int(vis.sum()), 26, 3))                                  # float64 (reference :119)
        kp[:, SMPL_TO_BODY26FK[:, 0]] = j2d[:, SMPL_TO_BODY26FK[:, 1]]
        d['kp_2d'], d['kp_2d_score'] = kp[:, :, :2], kp[:, :, 2]
        d['kp_2d_aligned'] = d['kp_2d'].copy()
        d['cam_K'] = est['cam_K'].astype(np.float32)
        if not np.all(visible):
            for key in ['kp_2d', 'kp_2d_score', 'kp_2d_aligned', 'cam_K']:
                full = np.zeros((n,) + d[key].shape[1:], dtype=d[key].dtype)
                full[vis] = d[key]
                d[key] = full
            vis_ind = np.where(visible)[0].astype(np.float32)
            for key in ['smpl_pose', 'smpl_beta', 'root_trans_cam', 'smpl_orient_cam']:
                f = interp1d(vis_ind, d[key], axis=0, assume_sorted=True, fill_value='extrapolate')
                d[key] = f(np.arange(n, dtype=np.float32))
        return _to_torch(d)

    def filter_pose(self, d):
        """:250-271 mark frames with an orientation jump > 60 deg as invisible."""
        visible = d['visible']
        q = rt.aa_to_quat(d['smpl_orient_cam'])
        jump = rt.quat_angle_diff(q[1:], q[:-1])
        ind = torch.where((jump > np.pi / 3) & visible[1:].bool())[0] + 1
        for i in ind:
            if visible[i - 1]:
                if i + 1 < q.shape[0] and visible[i + 1] and i + 1 not in ind:
                    visible[i - 1] = 0
                else:
                    visible[i] = 0
        if self.flag_make_invis_with_keypoint:
            vis_ind = torch.where(visible == 1.0)[0]
            nvalid = (d['kp_2d_score'][vis_ind] > self.make_invis_keypoint_min_score).sum(dim=1)
            visible[vis_ind[nvalid < self.make_invis_keypoint_min_num]] = 0.0
        d['vis_frames'] = visible == 1
        d['invis_frames'] = visible == 0

    def infer_motion_traj(self, d):
        """:353-392 (batch size 1 per person, sample_num 1)"""
        if self.mt_model is None:
            return
        ex = d['exist_frames']
        batch = {'in_body_pose': d['smpl_pose_nofill'][ex].unsqueeze(0).clone(),
                 'frame_mask': d['visible'][ex].unsqueeze(0).clone()}
        out = self.mt_model.inference(batch, sample_num=1)
        if self.flag_infill_motion:
            d['infilled'] = True
            d['smpl_pose'] = d['smpl_pose'].detach().clone()
            d['smpl_pose'][ex] = out['infer_out_body_pose'][0, 0].to(d['smpl_pose'].dtype)
        if self.flag_pred_traj:
            d['traj_predicted'] = True
            d['traj_local_pred'] = out['infer_out_local_traj_tp'][:, 0, 0, :].clone()
            d['smpl_orient_world_base'] = d['smpl_orient_world_base'].detach().clone()
            d['root_trans_world_base'] = d['root_trans_world_base'].detach().clone()
            if 'infer_out_pose' in out:
                d['smpl_orient_world_base'][ex] = out['infer_out_pose'][0, 0, :, :3]
            if 'infer_out_orient' in out:
                d['smpl_orient_world_base'][ex] = out['infer_out_orient'][0, 0]
            d['root_trans_world_base'][ex] = out['infer_out_trans'][0, 0]
            d['smpl_orient_world'] = d['smpl_orient_world_base']
            d['root_trans_world'] = d['root_trans_world_base']

    def init_default_traj(self, d):
        """:319-323"""
        d['root_trans_world_base'][:] = torch.tensor([0.0, 0.0, 0.8])
        d['smpl_orient_world_base'][:] = rt.quat_to_aa(torch.tensor([0.0, 0.0, 0.7071, 0.7071]))
        d['root_trans_world'] = d['root_trans_world_base']
        d['smpl_orient_world'] = d['smpl_orient_world_base']

    def init_cam_pose(self, data, all_frames=False):
        """:294-317 camera-to-world from the FIRST person's transform (the mean is commented out upstream)."""
        cands = [torch.matmul(d['person_transform_world'], d['person2cam']) * d['vis_frames'][:, None, None]
                 for d in data['person_data'].values()]
        npers = data['fr_num_persons']
        has = npers > 0
        start = torch.where(has)[0][0]
        inv = torch.zeros_like(data['cam_pose'])
        inv[has] = cands[0][has]
        data['pose_infer_cam_pose_inv'] = inv
        if all_frames:
            if not torch.all(has):
                last = inv[start]
                for i in range(len(npers)):
                    if npers[i] == 0:
                        data['cam_pose_inv'][i] = last
                    else:
                        last = data['cam_pose_inv'][i]
        else:
            inv[...] = inv[start].clone()
        inv[:, :3, :3] = rt.rot6d_to_rotmat(rt.rotmat_to_rot6d(inv[:, :3, :3]))
        data['cam_pose_inv'] = inv
        data['cam_pose'] = rt.inverse_transform(inv)

    def init_traj_heading_from_cam(self, data):
        """:273-292 overwrite the predicted heading of the cam-fixed frames with the camera-derived one."""
        for d in data['person_data'].values():
            world = torch.matmul(data['cam_pose_inv'], d['person_transform_cam'])
            q = rt.rotmat_to_quat(world[:, :3, :3].contiguous())
            q_interp = tc.interp_orient_q_sep_heading(q[d['vis_frames']], d['vis_frames'])
            local = tc.global_to_local(world[:, :3, 3], q_interp)
            for (s, e) in self.cam_fix_frames:
                d['traj_local_pred'][s:e, -2:] = local[d['exist_frames']][s:e, -2:]
            trans, oq = tc.local_to_global(d['traj_local_pred'])
            ex = d['exist_frames']
            d['smpl_orient_world_base'] = d['smpl_orient_world_base'].detach().clone()
            d['root_trans_world_base'] = d['root_trans_world_base'].detach().clone()
            d['smpl_orient_world_base'][ex] = rt.quat_to_aa(oq)
            d['root_trans_world_base'][ex] = trans
            d['smpl_orient_world'] = d['smpl_orient_world_base'].clone()
            d['root_trans_world'] = d['root_trans_world_base'].clone()
            d['person_transform_world'] = rt.make_transform(d['smpl_orient_world'], d['root_trans_world'], 'axis_angle')

    def init_data(self, in_dict):
        if self.est_type != 'hybrik':
            raise ValueError(f'est_type {self.est_type} not supported')
        num_fr = len(in_dict['est'][0]['bboxes_dict']['exist'])
        cam_pose = torch.eye(4).repeat(num_fr, 1, 1)
        cam_pose_inv = rt.inverse_transform(cam_pose)
        persons = {}
        for idx, est in in_dict['est'].items():
            d = self._person_from_estimate(est, in_dict['gt'].get(idx))
            if self.flag_filter_pose:
                self.filter_pose(d)
            d['root_trans_world'] = rt.transform_trans(cam_pose_inv, d['root_trans_cam'])
            d['smpl_orient_world'] = rt.transform_rot(cam_pose_inv, d['smpl_orient_cam'])
            d['root_trans_world_base'] = d['root_trans_world'].clone()
            d['smpl_orient_world_base'] = d['smpl_orient_world'].clone()
            d['smpl_pose_nofill'] = d['smpl_pose'].clone()
            d['smpl_pose_nofill'][~d['exist_frames']] = 0.0
            persons[idx] = d
        if self.flag_infer_motion_traj:
            for d in persons.values():
                self.infer_motion_traj(d)
        if not (self.flag_infer_motion_traj and self.flag_pred_traj):
            for d in persons.values():
                self.init_default_traj(d)
        for d in persons.values():
            d['person_transform_world'] = rt.make_transform(d['smpl_orient_world'], d['root_trans_world'], 'axis_angle')
            d['person_transform_cam'] = rt.make_transform(d['smpl_orient_cam'], d['root_trans_cam'], 'axis_angle')
            d['person2cam'] = rt.inverse_transform(d['person_transform_cam'])
        rel = None
        if self.flag_opt_traj:
            last = d                                                    # reference quirk: template = last person (:176)
            for d in persons.values():
                if self.flag_opt_person2cam_rot or self.flag_opt_person2cam_trans:
                    n = d['person2cam'].shape[0]
                    d['person2cam_res_rot'] = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(n, 1)
                    d['person2cam_res_trans'] = torch.zeros(n, 3)
                d['smpl_orient_world_res'] = torch.zeros_like(last['smpl_orient_world'])
                d['root_trans_world_res'] = torch.zeros_like(last['root_trans_world'])
            rel = {}
            ids = list(persons.keys())
            for i in range(len(ids)):
                for j in range(len(ids)):
                    if i != j:
                        rel[(i, j)] = torch.matmul(rt.inverse_transform(persons[ids[i]]['person_transform_cam']),
                                                   persons[ids[j]]['person_transform_cam'])
            if self.flag_pred_traj:
                for d in persons.values():
                    L = int(d['exist_len'].sum())
                    d['traj_local_xy'] = torch.zeros(2)
                    d['traj_local_dxy'] = torch.zeros(L - 1, 2)
                    d['traj_local_heading'] = torch.zeros(1)
                    d['traj_local_dheading'] = torch.zeros(L - 1)
                    d['traj_local_z'] = torch.zeros(L)
                    d['traj_local_rot'] = torch.zeros(L, 6)
            else:
                for d in persons.values():
                    d['root_trans_world_base'][:] = d['root_trans_world_base'][0].clone()
                    d['smpl_orient_world_base'][:] = d['smpl_orient_world_base'][0].clone()
        fr_num_persons = sum(d['vis_frames'] for d in persons.values())
        n_empty = int((fr_num_persons == 0).sum())
        data = {
            'seq_name': in_dict['seq_name'], 'person_data': persons, 'seq_len': num_fr,
            'fr_num_persons': fr_num_persons, 'cam_pose': cam_pose, 'cam_pose_inv': cam_pose_inv,
            'cam_inv_rot_residual': torch.zeros(n_empty, 6),
            'cam_inv_trans_residual': torch.zeros(num_fr if self.flag_cam_inv_trans_res_all else n_empty, 3),
            'rel_transform_cam': rel, 'gt': in_dict['gt'], 'gt_meta': in_dict['gt_meta'],
            'meta': {'algo': 'global_recon', 'num_fr': num_fr},
        }
        self.init_cam_pose(data)
        if self.flag_infer_motion_traj and self.flag_pred_traj:
            self.init_traj_heading_from_cam(data)
        if self.flag_init_cam_all_frames:
            self.init_cam_pose(data, all_frames=True)
        self.forward(data, [], {'stage': 'init'})
        return data

    # ------------------------------------------------------------------ forward (:394-531)
    def pred_trajectory_base(self, d):
        """:394-426 predicted local trajectory + delta variables -> world orientation / translation."""
        tl = d['traj_local_pred'].detach().clone()
        xy = torch.cat([(tl[0, :2] + d['traj_local_xy'])[None], tl[1:, :2] + d['traj_local_dxy']], dim=0)
        mask = torch.ones_like(tl[1:, 0])
        for (s, e) in self.cam_fix_frames:
            mask[s:e] = 0.0
        h0 = rt.vec_to_heading(tl[[0], -2:]) + d['traj_local_heading']
        hr = rt.vec_to_heading(tl[1:, -2:]) + d['traj_local_dheading'] * mask
        hvec = torch.cat([rt.heading_to_vec(h0), rt.heading_to_vec(hr)], dim=0)
        z = tl[:, 2] + d['traj_local_z']
        if self.flag_opt_vis_local_rot:
            vis = d['vis_frames'].to(tl.dtype)[:, None]
            d6 = tl[:, 3:-2] + d['traj_local_rot'] * vis
        else:
            d6 = tl[:, 3:-2] + d['traj_local_rot']
        d['traj_local'] = torch.cat([xy, z[:, None], d6, hvec], dim=-1)
        trans, oq = tc.local_to_global(d['traj_local'])
        ex = d['exist_frames']
        ob = d['smpl_orient_world_base'].detach().clone()
        tb = d['root_trans_world_base'].detach().clone()
        ob[ex] = rt.quat_to_aa(oq)
        tb[ex] = trans
        d['smpl_orient_world_base'], d['root_trans_world_base'] = ob, tb

    def forward(self, data, opt_variables, opt_meta):
        persons = data['person_data']
        for d in persons.values():
            if self.flag_infer_motion_traj and self.flag_pred_traj:
                self.pred_trajectory_base(d)
            if self.flag_opt_traj:
                if 'world_res' in opt_variables:
                    d['smpl_orient_world'] = d['smpl_orient_world_base'] + d['smpl_orient_world_res']
                    d['root_trans_world'] = d['root_trans_world_base'] + d['root_trans_world_res']
                else:
                    d['smpl_orient_world'] = d['smpl_orient_world_base']
                    d['root_trans_world'] = d['root_trans_world_base']
                if 'world_dheading' in d:
                    dh = d['world_dheading']
                    dq = rt.aa_to_quat(torch.cat([torch.zeros(dh.shape[0], 2, dtype=dh.dtype), dh], dim=-1))
                    d['smpl_orient_world'] = rt.quat_to_aa(rt.quat_mul(dq, rt.aa_to_quat(d['smpl_orient_world_base'])))
                    d['root_trans_world'] = d['root_trans_world_base']
                if 'world_dxy' in d:
                    raise NotImplementedError('world_dxy (in-place aliasing quirk, unused by shipped configs)')
            d['person_transform_world'] = rt.make_transform(d['smpl_orient_world'], d['root_trans_world'], 'axis_angle')

        if self.flag_opt_cam and opt_meta['stage'] != 'init':
            if 'cam' in opt_variables:
                T = data['cam_pose'].shape[0]
                if self.flag_fixed_cam:
                    data['cam_rot_6d'] = data['cam_rot_6d_fix'].expand(T, -1)
                    data['cam_trans'] = data['cam_trans_fix'].expand(T, -1)
                if 'cam_rot_6d' in data:
                    data['cam_pose'] = rt.make_transform(data['cam_rot_6d'], data['cam_trans'], '6d')
                    data['cam_pose_inv'] = rt.inverse_transform(data['cam_pose'])
            elif self.flag_opt_cam_from_person_pose:
                self._camera_from_persons(data)

        for d in persons.values():
            d['smpl_orient_cam_in_world'] = rt.transform_rot(data['cam_pose'], d['smpl_orient_world'])
            d['root_trans_cam_in_world'] = rt.transform_trans(data['cam_pose'], d['root_trans_world'])
            if 'smpl_pose' in d and 'cam_K' in d:
                dt = d['smpl_orient_world'].dtype                           # float32 (reference); float64 in the noise-floor runs
                joints, _ = self.smpl(d['smpl_orient_world'], d['smpl_pose'].to(dt),
                                      d['smpl_beta'].to(dt), root_trans=d['root_trans_world'],
                                      root_scale=d['scale'])
                d['joints_world'] = joints
                d['kp_2d_pred'] = rt.perspective_projection(rt.transform_trans(data['cam_pose'], joints), d['cam_K'])

    def _camera_from_persons(self, data):
        """:481-508 camera-to-world = mean over visible persons of person_transform_world @ person2cam, forward
        filled over frames without any person, plus 6d / translation residual variables."""
        cands = []
        for d in data['person_data'].values():
            p2c = d['person2cam']
            if self.flag_opt_person2cam_rot or self.flag_opt_person2cam_trans:
                p2c = torch.matmul(p2c, rt.make_transform(d['person2cam_res_rot'], d['person2cam_res_trans'], '6d'))
            cands.append(torch.matmul(d['person_transform_world'], p2c) * d['vis_frames'][:, None, None])
        npers = data['fr_num_persons']
        has = npers > 0
        mean = sum(cands) / npers.clamp(min=1)[:, None, None].to(cands[0].dtype)
        first = int(torch.where(has)[0][0])
        src = torch.zeros(len(npers), dtype=torch.long)                   # forward-fill index (:493-498)
        last = first
        for i in range(len(npers)):
            if npers[i] > 0:
                last = i
            src[i] = last
        inv = mean[src]
        d6 = rt.rotmat_to_rot6d(inv[:, :3, :3])
        empty = (npers == 0)
        if empty.any():
            add = torch.zeros_like(d6)
            add[empty] = data['cam_inv_rot_residual']
            d6 = d6 + add
        R = rt.rot6d_to_rotmat(d6)
        t = inv[:, :3, 3]
        if self.flag_cam_inv_trans_res_all:
            t = t + data['cam_inv_trans_residual']
        elif empty.any():
            add = torch.zeros_like(t)
            add[empty] = data['cam_inv_trans_residual']
            t = t + add
        data['cam_pose_inv'] = rt.make_transform(R, t)
        data['cam_pose'] = rt.inverse_transform(data['cam_pose_inv'])

    # ------------------------------------------------------------------ loss / optimiser (:533-644)
    def compute_loss(self, data, loss_cfg):
        total, weighted, unweighted = 0, {}, {}
        for name, specs in loss_cfg.items():
            if name not in RESIDUALS:
                raise KeyError(f'residual {name} not available')
            raw = RESIDUALS[name](data, specs)
            val = raw * specs['weight']
            if not specs.get('monitor_only', False):
                total = total + val
            weighted[name], unweighted[name] = val, raw
        return total, weighted, unweighted

    def get_parameter(self, data, opt_variables):
        """:591-633 (same ordering: camera block first, then per person in YAML key order)"""
        params = []
        if 'cam' not in opt_variables:
            params += [data['cam_inv_rot_residual'], data['cam_inv_trans_residual']]
        else:
            if self.flag_fixed_cam:
                data['cam_rot_6d_fix'] = rt.rotmat_to_rot6d(data['cam_pose'][[0], :3, :3]).detach()
                data['cam_trans_fix'] = data['cam_pose'][[0], :3, 3].clone().detach()
                params += [data['cam_rot_6d_fix'], data['cam_trans_fix']]
            else:
                data['cam_rot_6d'] = rt.rotmat_to_rot6d(data['cam_pose'][:, :3, :3]).detach()
                data['cam_trans'] = data['cam_pose'][:, :3, 3].clone().detach()
                params += [data['cam_rot_6d'], data['cam_trans']]
        for d in data['person_data'].values():
            if self.flag_opt_traj:
                for key in opt_variables:
                    if key == 'world_res':
                        params += [d['smpl_orient_world_res'], d['root_trans_world_res']]
                    if 'local' in key:
                        params.append(d[f'traj_{key}'])
            if self.flag_opt_person2cam_rot and 'person2cam_rot' in opt_variables:
                params.append(d['person2cam_res_rot'])
            if self.flag_opt_person2cam_trans and 'person2cam_trans' in opt_variables:
                params.append(d['person2cam_res_trans'])
            if 'world_dheading' in opt_variables:
                if 'world_dheading' not in d:
                    d['world_dheading'] = torch.zeros_like(d['smpl_orient_world'][..., [0]])
                params.append(d['world_dheading'])
            if 'world_dxy' in opt_variables:
                raise NotImplementedError('world_dxy')
        return params

    def optimize_main(self, data, opt_variables, opt_lr, opt_niters, loss_cfg, opt_meta, on_iter=None):
        params = self.get_parameter(data, opt_variables)
        for p in params:
            p.requires_grad_(True)
        opt = torch.optim.Adam(params, lr=opt_lr, betas=(0.9, 0.999)) if params else None
        last = {}

        def closure():
            opt.zero_grad()
            self.forward(data, opt_variables, opt_meta)
            loss, _, uw = self.compute_loss(data, loss_cfg)
            loss.backward()
            last['uw'], last['loss'] = uw, loss
            return loss

        for it in range(opt_niters):
            t0 = time.time()
            self.cur_iter = it
            if opt is not None:
                opt.step(closure)
            if on_iter is not None:
                on_iter(it, last, time.time() - t0)
        for p in params:
            p.requires_grad_(False)
        data['cam_pose'] = data['cam_pose'].detach()
        data['cam_pose_inv'] = data['cam_pose_inv'].detach()
        return data

    def optimize(self, in_dict, on_iter=None):
        data = self.init_data(in_dict)
        for stage, specs in self.opt_stage_specs.items():
            meta = {'stage': stage, 'opt_latent_start_iter': specs.get('opt_latent_start_iter', 0)}
            self.optimize_main(data, specs['opt_variables'], specs['opt_lr'], specs['opt_niters'], specs['loss_cfg'],
                               meta, on_iter)
            if specs.get('reinitialize_cam', False):
                data['cam_pose'][:] = data['cam_pose'][[0]]
                data['cam_pose_inv'] = rt.inverse_transform(data['cam_pose'])
        return _to_numpy(data)
