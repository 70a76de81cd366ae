"""Configuration objects with the attribute surface of the reference's ``Config`` classes.

``Config(cfg_id, out_dir)`` mirrors ``global_recon/utils/config.py:12-51`` (reference): ``.id``,
``.grecon_model_name``, ``.grecon_model_specs``, ``.opt_stage_specs``, ``.seed``, ``.cfg_dir``/``.log_dir``.  A
GLAMR YAML file is consumed unchanged: pass its path (or have it under ``global_recon/cfg/**/<id>.yml`` relative to
the cwd, as the reference expects).  The six shipped stage/weight tables (``global_recon/cfg/glamr_*.yml``) are also
available without any file through ``builtin_config_dict`` so that tests and benchmarks run where the reference tree
is absent; they are built from one base table plus per-config deltas.
"""
import copy
import glob
import os

import yaml


def _full_loss_cfg(rot_reg, inv_rot_smooth, origin_smooth, up_reg):
    return {
        'rel_transform': {'trans_weight': 0.0, 'weight': 200},
        'kp_2d': {'weight': 1.0, 'min_conf': 0.3},
        'kp_2d_dist': {'weight': 1.0, 'min_conf': 0.3, 'monitor_only': True},
        'cam_traj_rot': {'rot_type': '6d', 'weight': 1.0e5},
        'traj_rot_smoothness': {'weight': 1.0e3},
        'local_traj_dxy_reg': {'weight': 3.0e2},
        'local_traj_dheading_reg_new': {'weight': 3.0e3},
        'local_traj_rot_reg': {'weight': rot_reg},
        'local_traj_z_reg': {'weight': 1.0e2},
        'cam_inv_trans_residual_reg': {'weight': 1.0e2},
        'cam_inv_rot_smoothness': {'weight': inv_rot_smooth},
        'cam_origin_smoothness': {'weight': origin_smooth},
        'cam_up_reg': {'weight': up_reg},
    }


def _first_frame_loss_cfg():
    return {
        'rel_transform': {'trans_weight': 0.0, 'weight': 200, 'first_frame_only': True},
        'kp_2d': {'weight': 1.0, 'min_conf': 0.3, 'first_frame_only': True},
        'kp_2d_dist': {'weight': 1.0, 'min_conf': 0.3, 'monitor_only': True, 'first_frame_only': True},
        'cam_traj_rot': {'rot_type': '6d', 'weight': 1.0e5, 'first_frame_only': True},
    }


def _specs(dataset, **flags):
    s = {'motion_traj_cfg': 'joint_motion_traj_demo', 'est_type': 'hybrik', 'flag_infer_motion_traj': True,
         'flag_pred_traj': True, 'flag_opt_traj': True, 'flag_opt_cam': True}
    s.update(flags)
    return {'dataset': dataset, 'grecon_model_name': 'global_recon_model', 'grecon_model_specs': s}


_ALL_VARS = ['cam', 'local_xy', 'local_heading', 'world_dheading', 'local_dxy', 'local_rot', 'local_z']
_DYN_VARS = ['cam', 'local_xy', 'local_heading', 'world_dheading', 'local_rot']


def builtin_config_dict(cfg_id):
    """The YAML dict of a shipped config (global_recon/cfg/<cfg_id>.yml), rebuilt programmatically."""
    if cfg_id == 'glamr_dynamic':
        d = _specs('demo', flag_fixed_cam=False, flag_init_cam_all_frames=True)
        d['opt_stage_specs'] = {'init_opt': {'opt_lr': 1e-3, 'opt_niters': 500, 'opt_variables': list(_DYN_VARS),
                                             'loss_cfg': _full_loss_cfg(5e3, 1e1, 1e3, 1e6)}}
    elif cfg_id == 'glamr_static':
        d = _specs('demo', flag_fixed_cam=True)
        d['opt_stage_specs'] = {'init_opt': {'opt_lr': 1e-3, 'opt_niters': 500, 'opt_variables': list(_ALL_VARS),
                                             'loss_cfg': _full_loss_cfg(5e3, 1e3, 1e3, 1e2)}}
    elif cfg_id in ('glamr_static_multi', 'glamr_dynamic_multi'):
        static = cfg_id == 'glamr_static_multi'
        d = _specs('demo', flag_fixed_cam=True) if static else _specs('demo', flag_fixed_cam=False, flag_init_cam_all_frames=True)
        main = _full_loss_cfg(5e3, 1e3, 1e3, 1e2) if static else _full_loss_cfg(5e3, 1e1, 1e3, 1e6)
        d['opt_stage_specs'] = {
            'init_opt': {'opt_lr': 1e-1, 'opt_niters': 200, 'opt_variables': ['local_xy', 'local_heading'],
                         'loss_cfg': _first_frame_loss_cfg()},
            'main_opt': {'opt_lr': 1e-4, 'opt_niters': 500, 'opt_variables': list(_ALL_VARS if static else _DYN_VARS),
                         'loss_cfg': main}}
    elif cfg_id in ('glamr_3dpw', 'glamr_h36m'):
        pw = cfg_id == 'glamr_3dpw'
        d = _specs('3dpw' if pw else 'h36m', flag_fixed_cam=False, flag_init_cam_all_frames=False)
        if pw:
            d['grecon_model_specs']['flag_opt_cam_from_person_pose'] = True
        loss = _full_loss_cfg(5e2, 1e1, 1e2, 1e5) if pw else _full_loss_cfg(5e2, 1e4, 1e4, 1e5)
        v0 = ['local_xy', 'local_heading'] if pw else ['cam', 'local_xy', 'local_heading']
        v1 = (['local_xy', 'local_heading', 'local_dheading', 'local_dxy', 'local_rot'] if pw else
              ['cam', 'local_xy', 'local_heading', 'world_dheading', 'local_dxy', 'local_rot'])
        d['opt_stage_specs'] = {
            'init_opt': {'opt_lr': 1e-2, 'opt_niters': 200, 'opt_variables': v0, 'loss_cfg': copy.deepcopy(loss)},
            'main_opt': {'opt_lr': 1e-4, 'opt_niters': 500, 'opt_variables': v1, 'loss_cfg': copy.deepcopy(loss)}}
    else:
        raise KeyError(f'no built-in config named {cfg_id}')
    return d


BUILTIN_IDS = ['glamr_dynamic', 'glamr_static', 'glamr_static_multi', 'glamr_dynamic_multi', 'glamr_3dpw', 'glamr_h36m']


class Config:
    """Drop-in for ``global_recon.utils.config.Config``.

    cfg_id : a config id (resolved like the reference: ``global_recon/cfg/**/<id>.yml`` under the cwd, else the
             built-in table) or a path to a YAML file.
    """

    def __init__(self, cfg_id, out_dir=None, tmp=False, yml_dict=None):
        if yml_dict is not None:
            self.id, self.yml_file = cfg_id, None
        elif os.path.isfile(str(cfg_id)):
            self.yml_file = str(cfg_id)
            self.id = os.path.splitext(os.path.basename(self.yml_file))[0]
            yml_dict = yaml.safe_load(open(self.yml_file, 'r'))
        else:
            self.id = cfg_id
            files = glob.glob('global_recon/cfg/**/%s.yml' % cfg_id, recursive=True)
            if len(files) == 1:
                self.yml_file = files[0]
                yml_dict = yaml.safe_load(open(self.yml_file, 'r'))
            elif len(files) == 0:
                self.yml_file = None
                yml_dict = builtin_config_dict(cfg_id)
            else:
                raise AssertionError(f'ambiguous config id {cfg_id}: {files}')
        self.yml_dict = yml_dict
        if out_dir is None:
            root = os.path.expanduser(yml_dict.get('results_root_dir', 'results/global_recon'))
            self.cfg_dir = f'tmp/global_recon/{self.id}' if tmp else f'{root}/{self.id}'
        else:
            self.cfg_dir = out_dir
        self.log_dir = f'{self.cfg_dir}/logs'
        self.seed = yml_dict.get('seed', 1)
        self.grecon_model_name = yml_dict['grecon_model_name']
        self.grecon_model_specs = yml_dict.get('grecon_model_specs', dict())
        self.opt_stage_specs = yml_dict.get('opt_stage_specs', dict())
        for key in ['dataset', 'img_path', 'video_path', 'pose_path', 'bbox_path', 'pose_est_path', 'cam_est_path']:
            setattr(self, key, yml_dict.get(key, None))

    def make_dirs(self):
        os.makedirs(self.log_dir, exist_ok=True)

    def save_yml_file(self, out_path=None):
        out_path = out_path or f'{self.cfg_dir}/cfg.yml'
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, 'w') as f:
            yaml.safe_dump(self.yml_dict, f)
