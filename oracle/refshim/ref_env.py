"""Run the UNMODIFIED reference (/root/reference) in this container.

Test infrastructure only.  It exists so that (a) the oracle restatement under ``oracle/`` can be pinned against
the real reference and (b) ``tests/golden/make_golden.py`` can emit fixtures.  It needs ``/root/reference`` and
therefore never runs on the GPU box; nothing in the product, ``bench.py`` or the ``-m gpu`` tests imports it.

``activate()`` builds ``oracle/_ref/work`` (git-ignored):
    symlinks  global_recon lib motion_infiller traj_pred -> /root/reference/*
    data/body_models/smpl/SMPL_SYNTH.npz, data/J_regressor_extra.npy    (synthetic, glamr_b200.synthetic)
    results/.../checkpoints/model-best-epoch=0000.ckpt                  (empty files; shim ignores content)
then chdir()s there (the reference finds its YAML by cwd-relative glob) and puts the shims first on sys.path.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_ROOT = os.environ.get('GLAMR_REFERENCE_ROOT', '/root/reference')
WORK = os.path.join(REPO, 'oracle', '_ref', 'work')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'global_recon'))


def activate(asset_seed=0):
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    sys.path.insert(0, REPO)
    from glamr_b200.synthetic import make_smpl_assets
    os.makedirs(WORK, exist_ok=True)
    for d in ['global_recon', 'lib', 'motion_infiller', 'traj_pred']:
        dst = os.path.join(WORK, d)
        if not os.path.islink(dst):
            os.symlink(os.path.join(REF_ROOT, d), dst)
    smpl_dir = os.path.join(WORK, 'data', 'body_models', 'smpl')
    os.makedirs(smpl_dir, exist_ok=True)
    npz = os.path.join(smpl_dir, 'SMPL_SYNTH.npz')
    tag = os.path.join(smpl_dir, f'seed_{asset_seed}.tag')
    if not (os.path.exists(npz) and os.path.exists(tag)):
        a = make_smpl_assets(asset_seed)
        np.savez(npz, **{k: a[k] for k in ['v_template', 'shapedirs', 'posedirs', 'J_regressor', 'lbs_weights',
                                            'parents', 'faces']})
        np.save(os.path.join(WORK, 'data', 'J_regressor_extra.npy'), a['J_regressor_extra'])
        for f in os.listdir(smpl_dir):
            if f.endswith('.tag'):
                os.remove(os.path.join(smpl_dir, f))
        open(tag, 'w').close()
    for sub in ['motion_filler/motion_infiller_demo', 'traj_pred/traj_pred_demo']:
        cp = os.path.join(WORK, 'results', sub, 'version_0', 'checkpoints')
        os.makedirs(cp, exist_ok=True)
        open(os.path.join(cp, 'model-best-epoch=0000.ckpt'), 'a').close()
    os.chdir(WORK)
    for p in [WORK, os.path.join(HERE, 'shims')]:
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    return WORK


def make_reference_optimizer(cfg_id, niters=None):
    """-> (reference GlobalReconOptimizer on CPU, its Config).  `niters` overrides every stage's opt_niters."""
    import torch
    from global_recon.utils.config import Config
    from global_recon.models import model_dict
    cfg = Config(cfg_id, out_dir=os.path.join(WORK, 'out', cfg_id))
    if niters is not None:
        for st in cfg.opt_stage_specs.values():
            st['opt_niters'] = niters
    model = model_dict[cfg.grecon_model_name](cfg, torch.device('cpu'), None)
    return model, cfg
