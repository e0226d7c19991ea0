"""Wavelet model pre-processing and GWNN localisation on the device against the CPU
restatement (pygsp absent: parity unpinned, see oracle/__init__.py)."""
import numpy as np
import pytest
import torch

import oracle
from helpers import sym_graph


def test_oracle_spectral_preprocess_shapes_and_identity_limit():
    """tol = 0 and scale -> 0: both wavelets tend to the identity, so the localised block tends to relu(X)."""
    a = sym_graph(40, 120, 1)
    x = np.random.default_rng(0).standard_normal((40, 3)).astype(np.float32)
    lmax = oracle.estimate_lmax(oracle.combinatorial_laplacian(a))
    out = oracle.spectral_preprocess(a, x, 1e-9, 3, 0.0, lmax)
    assert out.shape == (40, 6)
    np.testing.assert_array_equal(out[:, :3], x)
    np.testing.assert_allclose(out[:, 3:], np.maximum(x, 0), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,block", [(300, 900, 1000), (700, 2500, 256)])
def test_spectral_model_preprocess_vs_oracle(n, m, block):
    from scalable_roubust_gnn_b200.spectral import SpectralModel
    a = sym_graph(n, m, 3)
    x = np.random.default_rng(1).standard_normal((n, 12)).astype(np.float32)
    lmax = oracle.estimate_lmax(oracle.combinatorial_laplacian(a))
    want, phis = oracle.spectral_preprocess(a, x, 0.5, 3, 1e-4, lmax, return_phis=True)
    model = SpectralModel(0.5, 3, 1e-4, lmax=lmax, block=block)
    got = model.preprocess(a, x)
    assert isinstance(got, torch.Tensor) and got.dtype == torch.float32 and tuple(got.shape) == (n, 24)
    np.testing.assert_array_equal(got[:, :12].numpy(), x)
    # (Psi Psi^-1) X is applied as Psi (Psi^-1 X): same map, other fp32 association
    np.testing.assert_allclose(got[:, 12:].numpy(), want[:, 12:], rtol=1e-5, atol=1e-6)
    for mine, ref in zip(model.phi_matrices, phis):
        ref.sort_indices()
        np.testing.assert_array_equal(mine.indptr, ref.indptr)
        np.testing.assert_array_equal(mine.indices, ref.indices)
        np.testing.assert_array_equal(mine.data, ref.data)
    dens = model.density()
    assert dens[0] == phis[0].nnz / n ** 2 and dens[1] == phis[1].nnz / n ** 2


@pytest.mark.gpu
def test_wavelet_localize_forward_backward_vs_dense():
    """GWNN layer core (wavelet/src/gwnn_layer.py:59-85, 111-128): Psi diag(theta) Psi^-1 (X W), gradients
    for theta, W and X, against dense fp64 torch on the CPU."""
    from scalable_roubust_gnn_b200.device import DeviceCSR
    from scalable_roubust_gnn_b200.sparse_mm import scipy_sparse_mat_to_device_adj
    from scalable_roubust_gnn_b200.spectral import wavelet_localize
    import scipy.sparse as sp
    n, f, c = 250, 9, 4
    rng = np.random.default_rng(2)
    phi = sp.random(n, n, 0.05, format="csr", dtype=np.float32, random_state=1)
    phi_inv = sp.random(n, n, 0.05, format="csr", dtype=np.float32, random_state=2)
    x = rng.standard_normal((n, f)).astype(np.float32)
    w = rng.standard_normal((f, c)).astype(np.float32)
    theta = (0.9 + 0.2 * rng.random((n, 1))).astype(np.float32)
    g = rng.standard_normal((n, c)).astype(np.float32)

    xd = torch.from_numpy(x).cuda().requires_grad_(True)
    wd = torch.from_numpy(w).cuda().requires_grad_(True)
    td = torch.from_numpy(theta).cuda().requires_grad_(True)
    out = wavelet_localize(scipy_sparse_mat_to_device_adj(phi), scipy_sparse_mat_to_device_adj(phi_inv),
                           torch.mm(xd, wd), theta=td)
    (out * torch.from_numpy(g).cuda()).sum().backward()

    P, Q = torch.from_numpy(phi.toarray()).double(), torch.from_numpy(phi_inv.toarray()).double()
    xr = torch.from_numpy(x).double().requires_grad_(True)
    wr = torch.from_numpy(w).double().requires_grad_(True)
    tr = torch.from_numpy(theta).double().requires_grad_(True)
    ref = P @ (tr * (Q @ (xr @ wr)))
    (ref * torch.from_numpy(g).double()).sum().backward()
    tol = dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), **tol)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), xr.grad.numpy(), **tol)
    np.testing.assert_allclose(wd.grad.cpu().numpy(), wr.grad.numpy(), **tol)
    np.testing.assert_allclose(td.grad.cpu().numpy(), tr.grad.numpy(), **tol)
