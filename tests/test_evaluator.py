"""Evaluator (global_recon/utils/evaluator.py): the oracle restatement against metrics produced by the executed reference
(tests/golden/evaluator.npz), the host-compiled Procrustes against the oracle's torch.svd formulation, and -- on the GPU --
the CUDA Evaluator against the same golden metrics."""
import copy
import ctypes

import numpy as np
import pytest
import torch

from helpers import load_golden

CASES = [('p2_t60', '3DPW', 2, 60, 250), ('p1_t300_h36m', 'h36m', 1, 300, 250), ('p1_t90_realign', '3DPW', 1, 90, 40)]
SCALARS = ['PA-MPJPE', 'PA-MPJPE-vis', 'PA-MPJPE-invis', 'G-MPJPE', 'G-MPVE', 'ACCEL']


def _case(gold, tag, P, T):
    from glamr_b200.synthetic import make_eval_case
    return make_eval_case(P, T, seed=int(gold[f'{tag}/seed']))


def _torchify(d):
    if isinstance(d, np.ndarray):
        return torch.tensor(d)
    if isinstance(d, dict):
        return {k: _torchify(v) for k, v in d.items()}
    return d


@pytest.mark.parametrize('tag,dataset,P,T,freq', CASES)
def test_oracle_evaluator_matches_reference(tag, dataset, P, T, freq, smpl_assets):
    from glamr_b200.synthetic import make_h36m_regressor
    from oracle.evaluator import OracleEvaluator
    gold = load_golden('evaluator')
    ev = OracleEvaluator(smpl_assets, make_h36m_regressor(0), dataset=dataset, align_freq=freq)
    got = ev.metrics(_torchify(_case(gold, tag, P, T)))
    for k in SCALARS:
        ref_val, ref_n = gold[f'{tag}/metric/{k}']
        assert got[k][1] == int(ref_n), (k, got[k][1], ref_n)
        assert abs(got[k][0] - ref_val) <= 1e-4 * max(abs(ref_val), 1.0), (k, got[k][0], ref_val)


def test_host_procrustes_matches_oracle():
    """the frame function of glamr_procrustes_align (host-compiled) vs lib/utils/torch_transform.py:282-345 restated with torch.svd,
    including a reflected and a planar (rank-deficient) frame"""
    import host_harness as hh
    from oracle.evaluator import similarity_align
    rng = np.random.default_rng(0)
    n, J = 64, 14
    S1 = rng.normal(0, 0.3, (n, J, 3)).astype(np.float32)
    S2 = (1.3 * S1[:, :, [1, 2, 0]] + rng.normal(0, 0.02, (n, J, 3)) + 0.5).astype(np.float32)
    S2[0] = S1[0] * np.array([1, 1, -1], np.float32)
    S1[1, :, 2] = 0
    S2[1, :, 2] = 0
    out = np.zeros_like(S1)
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    assert hh.lib().glamr_host_procrustes(n, J, fp(S1), fp(S2), fp(out)) == 0
    ref = similarity_align(torch.tensor(S1), torch.tensor(S2)).numpy()
    np.testing.assert_allclose(out, ref, atol=5e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('tag,dataset,P,T,freq', CASES)
def test_cuda_evaluator_matches_reference(tag, dataset, P, T, freq, smpl_assets):
    """compute_sequence_metrics on the GPU (4 SMPL passes with vertices, CSR H36M regression, Jacobi-SVD Procrustes) vs the
    metrics of the executed reference Evaluator on the same inputs"""
    import logging
    from glamr_b200.evaluator import Evaluator
    from glamr_b200.synthetic import make_h36m_regressor
    gold = load_golden('evaluator')
    ev = Evaluator('glamr', dataset, device=torch.device('cuda:0'), align_freq=freq, compute_sample=True, smpl=smpl_assets,
                   h36m_regressor=make_h36m_regressor(0), log=logging.getLogger('test_evaluator'))
    md = ev.compute_sequence_metrics(copy.deepcopy(_case(gold, tag, P, T)), 'case', accumulate=True)
    for k in SCALARS:
        ref_val, ref_n = gold[f'{tag}/metric/{k}']
        m = md['metrics'][k]
        assert m.count == int(ref_n), (k, m.count, ref_n)
        assert abs(m.avg - ref_val) <= 2e-4 * max(abs(ref_val), 1.0), (k, m.avg, ref_val)       # mm; 1e-4 m joints -> ~0.1 mm
    np.testing.assert_allclose(md['metrics']['sample_PA-MPJPE-invis'].avg, gold[f'{tag}/metric/sample_PA-MPJPE-invis'], atol=2e-2)
    line = ev.print_metrics(md, prefix='case --- ', print_accum=False)
    assert 'PA-MPJPE' in line and 'G-MPVE' in line
    seeds = ev.metrics_from_multiple_seeds([md, md])
    assert abs(seeds['metrics']['PA-MPJPE'].avg - md['metrics']['PA-MPJPE'].avg) < 1e-9
