"""Per-epoch sparse product of the GCN model: DeviceAdj behind torch.mm, forward and
backward, against one forward + backward of the reference's own Layer2GraphConvolution
(tests/golden/reference_ext.npz, made by tests/golden/make_golden_ext.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.nn as nn

import oracle
from conftest import GOLDEN as GOLDEN_DIR
from helpers import sym_graph


class Gcn2(nn.Module):
    """The node branch of Layer2GraphConvolution.forward (SSRG/models/base_scalable/simple_models.py:
    225-234), dropout off: fc1 -> adj -> relu -> fc2 -> adj.  `adj` is whatever torch.mm accepts."""

    def __init__(self, g, tag):
        super().__init__()
        self.fc1_node_edge = nn.Linear(20, 16)
        self.fc2_node = nn.Linear(16, 5)
        with torch.no_grad():
            for name, p in self.named_parameters():
                p.copy_(torch.from_numpy(g[f"{tag}_w_{name}"]))
        self.adj = None

    def forward(self, x):
        x = self.fc1_node_edge(x)
        x = torch.mm(self.adj, x)
        x = torch.relu(x)
        x = self.fc2_node(x)
        return torch.mm(self.adj, x)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))


def _norm_adj(r):
    return oracle.sym_norm(sym_graph(400, 2400, 6), r).tocsr()


@pytest.mark.parametrize("tag,r", [("gcn", 0.5), ("gcn_r03", 0.3)])
def test_mini_layer_restates_reference_layer_on_cpu(g, tag, r):
    """The test's Gcn2 + torch.sparse on the CPU reproduces the reference layer's recorded output, so
    the GPU test below compares like with like."""
    a = _norm_adj(r).tocoo().astype(np.float32)
    adj = torch.sparse_coo_tensor(np.vstack((a.row, a.col)).astype(np.int64), a.data, a.shape)
    net = Gcn2(g, tag)
    net.adj = adj
    y = net(torch.from_numpy(g[f"{tag}_x"]))
    np.testing.assert_allclose(y.detach().numpy(), g[f"{tag}_y"], rtol=1e-6, atol=1e-7)


def test_device_adj_contract():
    from scalable_roubust_gnn_b200.sparse_mm import scipy_sparse_mat_to_device_adj
    with pytest.raises(TypeError, match="scipy sparse matrix"):
        scipy_sparse_mat_to_device_adj(np.eye(3))
    with pytest.raises(ValueError, match="square"):
        scipy_sparse_mat_to_device_adj(sp.csr_matrix((2, 3)))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,r", [("gcn", 0.5), ("gcn_r03", 0.3)])
def test_gcn_forward_backward_vs_reference_golden(g, tag, r):
    from scalable_roubust_gnn_b200.operators import SymLaplacianGraphOp
    from scalable_roubust_gnn_b200.sparse_mm import scipy_sparse_mat_to_device_adj
    adj_n = SymLaplacianGraphOp(None, r=r).construct_adj(sym_graph(400, 2400, 6))      # device normalisation
    net = Gcn2(g, tag).cuda()
    net.adj = scipy_sparse_mat_to_device_adj(adj_n)
    x = torch.from_numpy(g[f"{tag}_x"]).cuda().requires_grad_(True)
    y = net(x)
    loss = torch.nn.functional.cross_entropy(y, torch.from_numpy(g[f"{tag}_target"]).cuda())
    loss.backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), g[f"{tag}_y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"][0], rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"{tag}_grad_x"], rtol=1e-4, atol=1e-7)
    for name, p in net.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), g[f"{tag}_g_{name}"], rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("f", [1, 7, 64, 100, 256])
def test_device_adj_mm_and_grad_vs_dense(f):
    from scalable_roubust_gnn_b200.sparse_mm import scipy_sparse_mat_to_device_adj
    rng = np.random.default_rng(f)
    a = sp.random(500, 500, 0.02, format="csr", dtype=np.float32, random_state=3)     # asymmetric, weighted
    a.sort_indices()
    adj = scipy_sparse_mat_to_device_adj(a)
    x = torch.from_numpy(rng.standard_normal((500, f)).astype(np.float32)).cuda().requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal((500, f)).astype(np.float32)).cuda()
    for fn in (torch.mm, torch.spmm, torch.sparse.mm, torch.matmul, lambda p, q: p @ q):
        x.grad = None
        y = fn(adj, x)
        (y * w).sum().backward()
        want_y = oracle.spmm_hop(a, x.detach().cpu().numpy())
        np.testing.assert_array_equal(y.detach().cpu().numpy(), want_y)                   # the hop kernel: bit-exact
        want_g = oracle.spmm_hop(a.T.tocsr(), w.cpu().numpy())
        np.testing.assert_array_equal(x.grad.cpu().numpy(), want_g)


@pytest.mark.gpu
def test_csr_transpose_exact_with_padding_and_empty_rows():
    from scalable_roubust_gnn_b200.device import DeviceCSR
    from scalable_roubust_gnn_b200.sparse_mm import csr_transpose
    a = sp.random(300, 300, 0.03, format="csr", dtype=np.float32, random_state=5).tolil()
    a[17, :] = 0
    a[:, 23] = 0
    a = a.tocsr()
    a.eliminate_zeros()
    a.sort_indices()
    pad = 100                                                       # arrays longer than the matrix
    ind = np.concatenate([a.indices, np.full(pad, 12345, np.int32)])
    dat = np.concatenate([a.data, np.full(pad, np.nan, np.float32)])
    d = DeviceCSR(torch.from_numpy(a.indptr).cuda(), torch.from_numpy(ind).cuda(), torch.from_numpy(dat).cuda(),
                  300, -1)
    t = csr_transpose(d)
    want = a.T.tocsr()
    want.sort_indices()
    m = want.nnz
    np.testing.assert_array_equal(t.indptr.cpu().numpy(), want.indptr)
    np.testing.assert_array_equal(t.indices.cpu().numpy()[:m], want.indices)
    np.testing.assert_array_equal(t.data.cpu().numpy()[:m], want.data)
    assert float(t.data[m:].abs().sum()) == 0.0


def test_device_adj_autograd_plumbing_on_cpu(monkeypatch):
    """torch.mm / spmm / sparse.mm / matmul / @ reach DeviceAdj through __torch_function__ and backward uses the
    transpose: checked on the CPU with the device hop replaced by a dense product (plumbing only)."""
    from scalable_roubust_gnn_b200 import sparse_mm as sm
    from scalable_roubust_gnn_b200.device import DeviceCSR
    a = sp.random(6, 6, 0.5, format="csr", dtype=np.float32, random_state=0)
    dense = torch.from_numpy(a.toarray())
    csr = DeviceCSR(torch.from_numpy(a.indptr), torch.from_numpy(a.indices), torch.from_numpy(a.data), 6, a.nnz)
    t_csr = DeviceCSR(csr.indptr, csr.indices, csr.data, 6, a.nnz)
    monkeypatch.setattr(sm.DeviceAdj, "_hop", staticmethod(lambda c, x: (dense if c is csr else dense.t()) @ x.detach()))
    adj = sm.DeviceAdj(csr)
    adj._t = sm.DeviceAdj(t_csr, transpose=adj)
    assert adj.t().t() is adj and tuple(adj.shape) == (6, 6) and adj.is_sparse
    for fn in (torch.mm, torch.spmm, torch.sparse.mm, torch.matmul, lambda p, q: p @ q):
        x = torch.rand(6, 3, requires_grad=True)
        y = fn(adj, x)
        y.sum().backward()
        assert torch.allclose(y, dense @ x) and torch.allclose(x.grad, dense.t() @ torch.ones(6, 3))
    with pytest.raises(TypeError):
        torch.add(adj, torch.ones(6, 6))          # anything but the products is not intercepted
