"""N>1 path on CPU: persons sharded over 2 ranks (gloo), one all-reduce of the packed [gradient | term sums] buffer,
must equal the single-rank result.  Uses the host-compiled frame functions (tests/host_harness) and the same
problem compiler / sharding fields (n_begin, n_end, owner) the CUDA path uses."""
import copy
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    from emu_runner import EmuRunner
    from glamr_b200.synthetic import make_smpl_assets
    from helpers import ReplayMT, case_setup
    from oracle.global_opt import OracleGlobalRecon
    assets = make_smpl_assets(0)
    gold, cfg, in_dict = case_setup(name, assets)
    results = {}
    for mode in ['single', 'sharded']:
        ora = OracleGlobalRecon(copy.deepcopy(cfg), assets, mt_model=ReplayMT(gold))
        data = ora.init_data(copy.deepcopy(in_dict))
        run = EmuRunner(ora, data)
        P = run.comp.P
        stage, specs = list(cfg.opt_stage_specs.items())[-1]
        if mode == 'single':
            run.set_stage(specs['opt_variables'], specs['loss_cfg'], stage)
        else:
            N = P * run.comp.T           # contiguous frame-person ranges; with P = 3 a person straddles the two ranks
            run.set_stage(specs['opt_variables'], specs['loss_cfg'], stage, n_begin=N * rank // world, n_end=N * (rank + 1) // world,
                          owner=(rank == 0))
        for it in range(3):
            run.backward()
            if mode == 'sharded':
                dist.all_reduce(run.reduce)          # the one collective per iteration
            run.step(specs['opt_lr'])
        results[mode] = (run.reduce.clone(), run.theta.clone())
    g_err = float((results['single'][0] - results['sharded'][0]).abs().max() / results['single'][0].abs().max())
    t_err = float((results['single'][1] - results['sharded'][1]).abs().max())
    ret[rank] = (g_err, t_err)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['static_multi_p3_t30', '3dpw_p2_t80_gaps'])
def test_person_sharding_allreduce_equals_single_rank(name):
    world, port = 2, _free_port()
    ret = mp.get_context('spawn').Manager().dict()
    mp.spawn(_worker, args=(world, port, name, ret), nprocs=world, join=True)
    for rank in range(world):
        g_err, t_err = ret[rank]
        assert g_err < 1e-5, f'rank {rank}: reduced gradient differs from single-rank by {g_err:.2e} (relative)'
        assert t_err < 1e-5, f'rank {rank}: parameters after 3 steps differ by {t_err:.2e}'
