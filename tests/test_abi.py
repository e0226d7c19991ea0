"""CPU suite: the C-ABI library loads, exports every symbol include/*.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls happen here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import ROOT
from scalable_roubust_gnn_b200 import _lib

HEADER = os.path.join(ROOT, "include", "srgnn_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:srg_\w+|FloatCSRMulDense\w*))\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/srgnn_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in _lib.SIGNATURES:
        assert re.search(rf"\bT {s}\b", exported), f"{s} bound in _lib.py but not a defined symbol"


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "srgnn_b200.h"\nint main(void){return srg_abi_version()==SRG_ABI_VERSION?0:1;}\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.srg_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    from scalable_roubust_gnn_b200.operators import SymLaplacianGraphOp, csr_sparse_dense_matmul
    assert _lib.device_count() == 0
    a = sp.identity(4, format="csr")
    x = np.ones((4, 3), np.float32)
    with pytest.raises(_lib.SrgError) as e:
        SymLaplacianGraphOp(2).propagate(a, x)
    assert e.value.code == _lib.SRG_ERR_NODEV
    with pytest.raises(_lib.SrgError):
        SymLaplacianGraphOp(2).construct_adj(a)
    with pytest.raises(_lib.SrgError):
        csr_sparse_dense_matmul(a, x)


def test_argument_validation_matches_reference_messages():
    from scalable_roubust_gnn_b200.operators import PprGraphOp, SymLaplacianGraphOp
    op = SymLaplacianGraphOp(2)
    x = np.ones((4, 3), np.float32)
    with pytest.raises(TypeError, match="The adjacency matrix must be a scipy csr sparse matrix!"):
        op.propagate(sp.identity(4, format="coo"), x)
    with pytest.raises(TypeError, match="The feature matrix must be a numpy.ndarray!"):
        op.propagate(sp.identity(4, format="csr"), [[1.0]])
    with pytest.raises(ValueError, match="Dimension mismatch detected for the adjacency and the feature matrix!"):
        op.propagate(sp.identity(5, format="csr"), x)
    with pytest.raises(ctypes.ArgumentError):
        op.propagate(sp.identity(4, format="csr"), x.astype(np.float64))
    with pytest.raises(TypeError, match="coo_matrix/csr_matrix"):
        PprGraphOp(2).construct_adj(sp.identity(4, format="csc"))
    assert SymLaplacianGraphOp(None).prop_steps is None       # gcn.py:8 builds the op without steps
    assert PprGraphOp(3).alpha == 0.15 and PprGraphOp(3).r == 0.5 and SymLaplacianGraphOp(3).r == 0.5


def test_host_all_ones_check():
    """Host utility of the unweighted-adjacency shortcut (on by default): multithreaded scan for values != 1."""
    lib = _lib.load()
    for dt, vt in ((np.float64, _lib.SRG_VAL_F64), (np.float32, _lib.SRG_VAL_F32)):
        a = np.ones(300_001, dtype=dt)
        assert lib.srg_host_all_ones(a.ctypes.data, vt, a.size, 8) == 1
        assert lib.srg_host_all_ones(a.ctypes.data, vt, 0, 8) == 1
        for pos in (0, 4095, 4096, a.size // 2, a.size - 1):
            b = a.copy()
            b[pos] = np.nextafter(dt(1), dt(2))
            for threads in (1, 3, 8):
                assert lib.srg_host_all_ones(b.ctypes.data, vt, b.size, threads) == 0, (pos, threads)
        z = a.copy()
        z[7] = 0                                   # an explicit zero is not "unweighted" either
        assert lib.srg_host_all_ones(z.ctypes.data, vt, z.size, 4) == 0
    assert lib.srg_host_all_ones(None, _lib.SRG_VAL_ONES, 10, 4) == 1
    assert lib.srg_host_all_ones(None, _lib.SRG_VAL_F64, 10, 4) == _lib.SRG_ERR_INVALID
