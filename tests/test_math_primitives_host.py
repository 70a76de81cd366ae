"""Forward values and hand-derived VJPs of glamr_b200/csrc/glamr_math.cuh (compiled for the host by
tests/host_harness) against the oracle's torch functions and torch autograd.  CPU only."""
import numpy as np
import pytest
import torch

import host_harness as hh
from oracle import rotations as rt
from oracle.smpl import rodrigues_smplx

OPS = {
    'aa_to_rotmat': (0, lambda a: rt.aa_to_rotmat(a).reshape(-1, 9)),
    'rodrigues_smplx': (1, lambda a: rodrigues_smplx(a).reshape(-1, 9)),
    'rot6d_to_rotmat': (2, lambda a: rt.rot6d_to_rotmat(a).reshape(-1, 9)),
    'rotmat_to_quat': (3, lambda a: rt.rotmat_to_quat(a.reshape(-1, 3, 3))),
    'quat_to_aa': (4, rt.quat_to_aa),
    'aa_to_quat': (5, rt.aa_to_quat),
    'rotmat_to_aa': (7, lambda a: rt.rotmat_to_aa(a.reshape(-1, 3, 3))),
}


def _inputs(name):
    g = torch.Generator().manual_seed(5)
    n = 256
    if name in ('aa_to_rotmat', 'aa_to_quat', 'rodrigues_smplx'):
        a = torch.randn(n, 3, generator=g)
        a[:8] *= 1e-4            # theta^2 < 1e-6 (Taylor / clamped branch)
        a[8:16] *= 2e-3
        a[16:24] = torch.nn.functional.normalize(a[16:24], dim=-1) * 3.1
        if name != 'rodrigues_smplx':
            a[24:28] = 0.0       # exactly on the branch point: gradient of the small-angle expression
        return a
    if name == 'rot6d_to_rotmat':
        return torch.randn(n, 6, generator=g)
    if name in ('rotmat_to_quat', 'rotmat_to_aa'):
        aa = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * torch.linspace(0.01, 3.12, n)[:, None]
        R = rt.aa_to_rotmat(aa) + 1e-3 * torch.randn(n, 3, 3, generator=g)   # slightly off SO(3): all 9 partials matter
        return R.reshape(n, 9)
    if name == 'quat_to_aa':
        q = torch.randn(n, 4, generator=g)
        q[:64] = torch.nn.functional.normalize(q[:64], dim=-1)
        q[64] = torch.tensor([1.0, 0.0, 0.0, 0.0])
        q[65] = torch.tensor([-1.0, 0.0, 0.0, 0.0])
        q[66] = torch.tensor([0.9999999, 3e-4, 0.0, 0.0])      # s^2 < 1e-6: clamp active, zero grad through s
        q[67] = torch.tensor([-0.5, 0.5, 0.5, 0.5])
        return q
    raise KeyError(name)


@pytest.mark.parametrize('name', list(OPS))
def test_forward_and_vjp(name):
    op, fn = OPS[name]
    a = _inputs(name).clone().requires_grad_(True)
    ref = fn(a)
    out = hh.rowop_fwd(op, a.detach().numpy())
    np.testing.assert_allclose(out, ref.detach().numpy(), atol=3e-6, err_msg=f'{name} forward')
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
    (gref,) = torch.autograd.grad(ref, a, gout)
    ga, _ = hh.rowop_vjp(op, a.detach().numpy(), None, gout.numpy())
    scale = np.maximum(np.abs(gref.numpy()).max(axis=1, keepdims=True), 1.0)
    err = np.abs(ga - gref.numpy()) / scale
    assert err.max() < 2e-4, f'{name} vjp: max rel err {err.max():.3e} at row {err.max(axis=1).argmax()}'


def test_quat_mul_and_mat3():
    g = torch.Generator().manual_seed(2)
    a = torch.randn(64, 4, generator=g, requires_grad=True)
    b = torch.randn(64, 4, generator=g, requires_grad=True)
    ref = rt.quat_mul(a, b)
    np.testing.assert_allclose(hh.rowop_fwd(6, a.detach().numpy(), b.detach().numpy()), ref.detach().numpy(), atol=2e-6)
    go = torch.randn(64, 4, generator=g)
    gra, grb = torch.autograd.grad(ref, (a, b), go)
    ga, gb = hh.rowop_vjp(6, a.detach().numpy(), b.detach().numpy(), go.numpy(), want_b=True)
    np.testing.assert_allclose(ga, gra.numpy(), atol=1e-5)
    np.testing.assert_allclose(gb, grb.numpy(), atol=1e-5)
    A = torch.randn(32, 3, 3, generator=g, requires_grad=True)
    B = torch.randn(32, 3, 3, generator=g, requires_grad=True)
    ref = torch.matmul(A, B)
    go = torch.randn(32, 3, 3, generator=g)
    gra, grb = torch.autograd.grad(ref, (A, B), go)
    ga, gb = hh.rowop_vjp(11, A.detach().numpy(), B.detach().numpy(), go.numpy(), want_b=True)
    np.testing.assert_allclose(ga, gra.reshape(-1, 9).numpy(), atol=1e-5)
    np.testing.assert_allclose(gb, grb.reshape(-1, 9).numpy(), atol=1e-5)


def test_safe_atan2_and_project():
    g = torch.Generator().manual_seed(4)
    yx = torch.randn(128, 2, generator=g)
    yx[:4] = 0.0
    yx[4:8] = torch.tensor([1e-7, -1e-7])
    yx = yx.requires_grad_(True)
    ref = rt.safe_atan2(yx[:, 0], yx[:, 1])
    np.testing.assert_allclose(hh.rowop_fwd(9, yx.detach().numpy())[:, 0], ref.detach().numpy(), atol=1e-6)
    go = torch.randn(128, generator=g)
    (gr,) = torch.autograd.grad(ref, yx, go)
    ga, _ = hh.rowop_vjp(9, yx.detach().numpy(), None, go.numpy()[:, None])
    np.testing.assert_allclose(ga, gr.numpy(), rtol=1e-4, atol=1e-4 * np.abs(gr.numpy()).max())
    X = (torch.randn(64, 3, generator=g) + torch.tensor([0.0, 0.0, 5.0])).requires_grad_(True)
    K = torch.tensor([[1000.0, 0, 960], [0, 1000.0, 540], [0, 0, 1]]).repeat(64, 1, 1)
    ref = rt.perspective_projection(X[:, None], K)[:, 0]
    np.testing.assert_allclose(hh.rowop_fwd(10, X.detach().numpy(), K.reshape(-1, 9).numpy()), ref.detach().numpy(), rtol=1e-6, atol=1e-3)
    go = torch.randn(64, 2, generator=g)
    (gr,) = torch.autograd.grad(ref, X, go)
    ga, _ = hh.rowop_vjp(10, X.detach().numpy(), K.reshape(-1, 9).numpy(), go.numpy())
    np.testing.assert_allclose(ga, gr.numpy(), rtol=1e-4, atol=1e-2)
