"""Sequence sweep driver (reference: global_recon/run_dataset.py:60-112, run_demo.py:55-82): sharding of independent
sequences over ranks, file naming, pickle round trip.  CPU part uses a stub optimiser; the GPU part runs the real one."""
import os
import pickle

import numpy as np
import pytest

from glamr_b200.global_recon import run_dataset as rd


class _StubModel:
    def __init__(self):
        self.seen = []

    def optimize(self, in_dict):
        self.seen.append(in_dict['seq_name'])
        return {'seq_name': in_dict['seq_name'], 'n_est': len(in_dict['est']), 'gt': in_dict['gt'], 'rand': float(np.random.rand())}


def test_shard_is_a_partition():
    items = [f's{i}' for i in range(11)]
    for world in (1, 2, 3, 8):
        parts = [rd.shard(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(items)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_sweep_reads_pose_pkl_and_writes_reference_layout(tmp_path, monkeypatch):
    pose_root = tmp_path / 'poses'
    for i, name in enumerate(['seqA', 'seqB', 'seqC']):
        if i % 2 == 0:                                  # both on-disk layouts
            d = pose_root / name / 'pose_est'
            d.mkdir(parents=True)
            f = d / 'pose.pkl'
        else:
            pose_root.mkdir(exist_ok=True)
            f = pose_root / f'{name}.pkl'
        with open(f, 'wb') as fh:
            pickle.dump({0: {'tag': name}, 1: {'tag': name}}, fh)
    out_dir = tmp_path / 'out'
    seen = {}
    for rank in range(2):
        monkeypatch.setenv('RANK', str(rank))
        monkeypatch.setenv('WORLD_SIZE', '2')
        args = rd.parse(['--cfg', 'glamr_3dpw', '--out_dir', str(out_dir), '--pose_root', str(pose_root), '--seeds', '1,7', '--quiet'])
        stub = _StubModel()
        done = rd.run(args, make_model=lambda cfg, local: stub)
        seen[rank] = stub.seen
        for seq, seed, path, _ in done:
            assert path == os.path.join(str(out_dir), seq, 'grecon', f'{seq}_seed{seed}.pkl')     # run_dataset.py:93
            with open(path, 'rb') as fh:
                out = pickle.load(fh)
            assert out['seq_name'] == seq and out['n_est'] == 2 and out['gt'] == {}
    assert sorted(set(seen[0])) == ['seqA', 'seqC'] and sorted(set(seen[1])) == ['seqB']          # round-robin replicas
    assert len(seen[0]) == 4 and len(seen[1]) == 2                                                 # two seeds each
    # the seed is set before each call (run_dataset.py:85-86): same seed -> same draw
    a = pickle.load(open(rd.out_file_of(str(out_dir), 'seqA', 7), 'rb'))['rand']
    c = pickle.load(open(rd.out_file_of(str(out_dir), 'seqC', 7), 'rb'))['rand']
    assert a == c
    # --cached 1 skips existing outputs
    monkeypatch.setenv('RANK', '0')
    args = rd.parse(['--out_dir', str(out_dir), '--pose_root', str(pose_root), '--seeds', '1,7', '--cached', '1', '--quiet'])
    stub = _StubModel()
    rd.run(args, make_model=lambda cfg, local: stub)
    assert stub.seen == []


def test_missing_pose_file_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        rd.find_pose_file(str(tmp_path), 'nope')


@pytest.mark.gpu
def test_synthetic_sweep_matches_direct_calls(tmp_path):
    import copy
    import torch
    from glamr_b200.config import Config
    from glamr_b200.global_recon.models import model_dict
    from glamr_b200.motion_traj import MotionTrajJointModel
    from glamr_b200.smpl import SMPL
    from glamr_b200.synthetic import make_in_dict, make_smpl_assets
    from glamr_b200.synthetic_nets import make_prior_states
    args = rd.parse(['--cfg', 'glamr_3dpw', '--out_dir', str(tmp_path), '--synthetic', '2', '--frames', '64', '--gaps', '--quiet'])
    done = rd.run(args)
    assert [d[0] for d in done] == ['synthetic_0000', 'synthetic_0001']
    dev = torch.device('cuda', 0)
    assets = make_smpl_assets(0)
    smpl = SMPL(assets, device=dev)
    cfg = Config('glamr_3dpw', out_dir=str(tmp_path))
    model = model_dict[cfg.grecon_model_name](cfg, dev, None, smpl=smpl,
                                              mt_model=MotionTrajJointModel(None, dev, None, smpl=smpl, states=make_prior_states(1234)))
    for i, (seq, seed, path, _) in enumerate(done):
        np.random.seed(seed)
        torch.manual_seed(seed)
        ref = model.optimize(make_in_dict(assets, 1, 64, seed=i, gaps=True, seq_name=seq))
        out = pickle.load(open(path, 'rb'))
        for k in ['smpl_orient_world', 'root_trans_world', 'smpl_pose']:
            np.testing.assert_array_equal(out['person_data'][0][k], ref['person_data'][0][k])    # same kernels, same seeds: bit-identical
        np.testing.assert_array_equal(out['cam_pose'], ref['cam_pose'])


def test_run_demo_layout(tmp_path):
    from glamr_b200.global_recon import run_demo
    pose_dir = tmp_path / 'pose_est'
    pose_dir.mkdir()
    with open(pose_dir / 'pose.pkl', 'wb') as fh:
        pickle.dump({0: {'tag': 'x'}}, fh)
    stub = _StubModel()
    argv = ['--cfg', 'glamr_static', '--video_path', 'assets/static/basketball.mp4', '--out_dir', str(tmp_path), '--seed', '3']
    out = run_demo.main(argv, make_model=lambda cfg, gpu: stub)
    assert out['seq_name'] == 'basketball' and stub.seen == ['basketball']
    assert os.path.exists(tmp_path / 'grecon' / 'basketball_seed3.pkl')                       # run_demo.py:74
    run_demo.main(argv, make_model=lambda cfg, gpu: stub)                                      # cached: no second call
    assert stub.seen == ['basketball']
    with pytest.raises(FileNotFoundError):
        run_demo.main(['--out_dir', str(tmp_path / 'other'), '--cached', '0'], make_model=lambda cfg, gpu: stub)
