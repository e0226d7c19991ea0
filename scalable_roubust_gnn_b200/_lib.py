"""ctypes binding of libsrgnn_b200.so (the C ABI declared in include/srgnn_b200.h).

The library is the product: there is no Python / CPU fallback behind these calls.  If the shared
object has not been built the import of any compute entry point fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC_DIR, "libsrgnn_b200.so")

# error codes (include/srgnn_b200.h)
SRG_OK = 0
SRG_ERR_INVALID = -22
SRG_ERR_NOMEM = -12
SRG_ERR_CUDA = -5
SRG_ERR_NODEV = -19
SRG_ERR_RANGE = -34
SRG_ERR_UNSUPPORTED = -95

SRG_VAL_ONES, SRG_VAL_F32, SRG_VAL_F64 = 0, 1, 2
SRG_FLAG_UNSORTED, SRG_FLAG_ASYMMETRIC, SRG_FLAG_ZERO_PRODUCT, SRG_FLAG_BAD_INDEX, SRG_FLAG_WEIGHTED = 1, 2, 4, 8, 16
SRG_FLAG_EXPLICIT_ZERO = 32
SRG_VAL_HAS_ZEROS = 0x100
SRG_AGG_NONE, SRG_AGG_LAST, SRG_AGG_SUM, SRG_AGG_MEAN, SRG_AGG_MAX, SRG_AGG_MIN, SRG_AGG_CONCAT, SRG_AGG_WEIGHTED = range(8)
SRG_AGG_NAFS = 8


class SrgError(RuntimeError):
    """A libsrgnn_b200 call failed; ``code`` is the negative errno-style return value."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libsrgnn_b200 error {code}: {message}")
        self.code = code
        self.message = message


class SrgUnsupported(SrgError):
    """The input needs a path the device library does not implement (message says which)."""


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/srgnn_b200.h declares
SIGNATURES = {
    "srg_abi_version": (C.c_int, []),
    "srg_last_error": (C.c_char_p, []),
    "srg_device_count": (C.c_int, []),
    "srg_launch_count": (_i64, []),
    "srg_host_all_ones": (C.c_int, [_vp, C.c_int, _i64, _i32]),
    "srg_set_tuning": (C.c_int, [C.c_char_p, _i64]),
    "srg_degree_selfloop_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "srg_sym_norm_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "srg_selfloop_rows_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "srg_selfloop_fill_rows_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "srg_pow_tables_f64": (C.c_int, [_vp, _i64, _f64, _vp, _vp, _vp]),
    "srg_norm_values_rows_csr": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _f64, C.c_int, _vp, _vp, _vp, _vp]),
    "srg_sym_norm_csr_general": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "srg_csr_canonicalize": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "srg_mag_norm_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _f64, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp]),
    "srg_spgemm_csr_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _i64,
                                     C.POINTER(_i64), _vp, _vp]),
    "srg_csr_append_diagonal": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "srg_csr_sym_scale_f32": (C.c_int, [_vp, _vp, _vp, _i64, C.c_float, _vp, _vp, _vp]),
    "srg_ppr_iterate_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp, _vp]),
    "srg_ppr_symmetrize": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "srg_teleport_iterate_f64": (C.c_int, [_vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp, _vp]),
    "srg_csr_intersect_mean_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "srg_synth_rmat_shard_csr": (C.c_int, [C.c_uint64, _i32, _i64, _f64, _f64, _f64, _i64, _i64, _i64, _i64, _vp, _vp,
                                           C.POINTER(_i64), _vp]),
    "srg_synth_hash_features_f32": (C.c_int, [C.c_uint64, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp]),
    "srg_csr_transpose_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "srg_edge_gather_i64": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "srg_edges_to_sym_csr": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "srg_apply_feature_mask_f32": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _vp]),
    "srg_spmm_csr_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _i32, _vp]),
    "srg_spmm_csr_f32_push": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, C.POINTER(_vp), _i32, _i64, _i64, _i32, _vp]),
    "srg_spmm_csr_f32_push2": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, C.POINTER(_vp), C.POINTER(_i64), _i32, _i64,
                                         _i32, _vp]),
    "srg_peer_barrier": (C.c_int, [_vp, C.POINTER(_vp), _i32, _i32, C.c_uint32, _vp, _vp]),
    "srg_push_rows_f32": (C.c_int, [_vp, _i64, _i64, C.POINTER(_vp), _i32, _i64, _vp]),
    "srg_dist_unique_id": (C.c_int, [_vp]),
    "srg_dist_init": (C.c_int, [_vp, _i32, _i32, _i64, _i32, C.POINTER(_vp)]),
    "srg_dist_init_comm": (C.c_int, [_vp, _i32, _i32, _i64, _i32, C.POINTER(_vp)]),
    "srg_dist_partition": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "srg_dist_propagate": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i64, _vp, _i64, _i32, _f64, _f64, C.POINTER(_vp), _i64,
                                     _vp, _vp]),
    "srg_dist_destroy": (C.c_int, [_vp]),
    "srg_copy_async": (C.c_int, [_vp, _vp, _i64, _vp]),
    "srg_ipc_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "srg_ipc_free": (C.c_int, [_vp]),
    "srg_ipc_get_handle": (C.c_int, [_vp, _vp]),
    "srg_ipc_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "srg_ipc_close": (C.c_int, [_vp]),
    "srg_propagate_khop_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.POINTER(_vp), _i64, _i32, _i32, _vp]),
    "srg_laplacian_csr": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "srg_cheby_filter_f64": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i32, _f64, C.POINTER(_f64), _i32, _i32, _f64,
                                       C.POINTER(_vp), C.POINTER(_vp), _i64, _vp, _vp, _vp]),
    "srg_cheby_sparse_run": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _f64, C.POINTER(_f64), _i32, _i32, _f64, C.POINTER(_vp), _vp]),
    "srg_cheby_sparse_info": (C.c_int, [_vp, _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "srg_cheby_sparse_fetch": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "srg_cheby_sparse_free": (C.c_int, [_vp]),
    "srg_endpoint_counts_i64": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "srg_candidate_topk_f32": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "srg_csr_to_edge_index_i64": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "srg_lanczos_lambda_max_f64": (C.c_int, [_vp, _vp, _vp, _i64, _f64, _i32, C.POINTER(_f64), C.POINTER(_i32), _vp]),
    "srg_pack_features_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp]),
    "srg_unpack_features_f32": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "srg_dense_block_to_csr_f32": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "srg_csr_block_scatter_f32": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "srg_csr_row_normalize_l1_f32": (C.c_int, [_i64, _vp, _vp, _vp]),
    "srg_exclusive_scan_i32": (C.c_int, [_vp, _i64, _vp, _vp]),
    "srg_aggregate_update_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i64, _i32, _i32, C.c_float, _i32, _vp]),
    "srg_nafs_combine_f32": (C.c_int, [C.POINTER(_vp), _i32, _i64, _i64, _i32, _vp, _i64, _vp, _vp]),
    "srg_propagate_aggregate_host": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _vp, _i32, _vp, _i32, _f64, _f64,
                                               _i32, _i32, _i32, _vp, _vp, C.c_int]),
    "srg_propagate_host": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _vp, _i32, _vp, _i32, _f64, _f64,
                                     C.POINTER(_vp), _vp, _vp, _vp, C.POINTER(_i64), C.c_int]),
    "srg_construct_adj_host": (C.c_int, [_vp, _vp, _vp, C.c_int, _i64, _i64, _f64, _f64, _vp, _vp, _vp,
                                         C.POINTER(_i64), C.c_int]),
    "srg_release_workspace": (C.c_int, []),
    "FloatCSRMulDenseOMP": (None, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "FloatCSRMulDense": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
}

_lib = None
_lock = threading.Lock()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into libsrgnn_b200.so (in-tree)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libsrgnn_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (never builds implicitly; never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C scalable_roubust_gnn_b200/csrc`. There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
                fn.restype = res
                fn.argtypes = args
            if lib.srg_abi_version() != 1:
                raise RuntimeError("libsrgnn_b200.so ABI version mismatch")
            # experiment knobs from the environment: SRG_TUNE="bulk_tile=1,bulk_rows=8" (srg_set_tuning keys)
            for item in filter(None, os.environ.get("SRG_TUNE", "").split(",")):
                key, _, val = item.partition("=")
                if lib.srg_set_tuning(key.strip().encode(), int(val)) != SRG_OK:
                    raise RuntimeError(f"SRG_TUNE: {lib.srg_last_error().decode()}")
            _lib = lib
    return _lib


def last_error() -> str:
    return load().srg_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == SRG_OK:
        return
    msg = last_error()
    if rc == SRG_ERR_UNSUPPORTED:
        raise SrgUnsupported(rc, msg)
    if rc == SRG_ERR_INVALID:
        raise SrgError(rc, msg)
    raise SrgError(rc, msg)


def device_count() -> int:
    return int(load().srg_device_count())


def launch_count() -> int:
    return int(load().srg_launch_count())


def set_tuning(key: str, value: int) -> None:
    check(load().srg_set_tuning(key.encode(), int(value)))
