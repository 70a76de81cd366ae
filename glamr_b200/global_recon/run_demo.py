"""Single-video entry point: what ``global_recon/run_demo.py:55-82`` does after pose estimation --
``<pose_est_dir>/pose.pkl`` -> ``GlobalReconOptimizer.optimize`` -> ``<out_dir>/grecon/<seq>_seed<seed>.pkl``.

Pose estimation and rendering are out of scope: run the reference's ``run_pose_est_on_video`` first, or
point ``--pose_est_dir`` at an existing HybrIK ``pose.pkl``; the pickle written here is the file the reference's
``GReconVisualizer`` loads (``run_demo.py:84-99``).

    python -m glamr_b200.global_recon.run_demo --cfg glamr_static --video_path assets/static/basketball.mp4 \
        --out_dir out/glamr_static/basketball [--pose_est_dir ...] [--seed 1] [--gpu 0] [--cached 1]
"""
import argparse
import os
import pickle

import numpy as np


def main(argv=None, make_model=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--cfg', default='glamr_static')
    ap.add_argument('--video_path', default='assets/static/basketball.mp4')
    ap.add_argument('--out_dir', default='out/glamr_static/basketball')
    ap.add_argument('--pose_est_dir', default=None)
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--gpu', type=int, default=0)
    ap.add_argument('--cached', type=int, default=1)
    args = ap.parse_args(argv)

    from glamr_b200.config import Config
    cfg = Config(args.cfg, out_dir=args.out_dir)
    seq_name = os.path.splitext(os.path.basename(args.video_path))[0]                # run_demo.py:43-46
    pose_est_dir = args.pose_est_dir or os.path.join(args.out_dir, 'pose_est')
    pose_est_file = os.path.join(pose_est_dir, 'pose.pkl')
    grecon_path = os.path.join(args.out_dir, 'grecon')
    os.makedirs(grecon_path, exist_ok=True)
    out_file = os.path.join(grecon_path, f'{seq_name}_seed{args.seed}.pkl')          # run_demo.py:74
    if args.cached and os.path.exists(out_file):
        with open(out_file, 'rb') as f:
            return pickle.load(f)
    if not os.path.exists(pose_est_file):
        raise FileNotFoundError(f'{pose_est_file} not found: pose estimation (HybrIK) is not part of glamr_b200')
    if make_model is None:
        import torch
        from glamr_b200.global_recon.models import model_dict
        device = torch.device('cuda', args.gpu)
        torch.cuda.set_device(device)
        model = model_dict[cfg.grecon_model_name](cfg, device, None)
        torch.manual_seed(args.seed)
    else:
        model = make_model(cfg, args.gpu)
    np.random.seed(args.seed)
    with open(pose_est_file, 'rb') as f:
        est_dict = pickle.load(f)
    in_dict = {'est': est_dict, 'gt': dict(), 'gt_meta': dict(), 'seq_name': seq_name}   # run_demo.py:78-79
    out_dict = model.optimize(in_dict)
    with open(out_file, 'wb') as f:
        pickle.dump(out_dict, f)
    return out_dict


if __name__ == '__main__':
    main()
