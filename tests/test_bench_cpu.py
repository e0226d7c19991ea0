"""CPU-side checks of bench.py: the reference arm (the unmodified reference from baseline/_ref on a small shape), the
byte accounting of the roofline block, and the default multi-GPU grid."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_reference(workload, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line_on_cora_shape():
    d = _run_reference("cora", {"OMP_NUM_THREADS": "1"})      # torch.distributed.run exports exactly this
    assert d["impl"] == "reference" and d["unit"] == "edge*feat/s" and d["higher_is_better"] is True
    assert d["config"]["N"] == 2708 and d["config"]["F"] == 1433 and d["config"]["K"] == 3
    assert d["config"]["nnz_hat"] == int(d["config"]["nnz_hat"]) > 2708
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the OpenMP thread count pinned to 1 by the launcher is restored to the host's CPU count
    assert cb["omp_max_threads"] == (os.cpu_count() or 1) and cb["cores"] == cb["omp_max_threads"]
    from baseline import ref_arm
    if ref_arm.available():
        assert cb["kind"] == "reference" and "UNMODIFIED reference" in cb["sample"]
        assert d["executed_steps"] == 2 and cb["norm_s"] > 0 and cb["hop_s"] > 0
    else:
        assert cb["kind"] in ("port", "reference")


def test_roofline_byte_accounting_matches_survey_8d():
    sys.path.insert(0, ROOT)
    import bench
    n, nnz_hat, f = 2449029, 64307827, 100
    assert bench.gather_bytes(n, nnz_hat, f) == nnz_hat * 8 + (n + 1) * 4 + nnz_hat * f * 4 + n * f * 4 == 27227001136
    assert bench.comp_bytes(n, nnz_hat, f) == nnz_hat * 8 + (n + 1) * 4 + 2 * n * f * 4


def test_default_grid_of_the_multi_gpu_bench():
    from scalable_roubust_gnn_b200 import dist as sdist
    from scalable_roubust_gnn_b200.dist_bench import default_feat_groups
    assert [default_feat_groups(w, "push") for w in (2, 4, 8)] == ["1", "1", "2"]
    assert default_feat_groups(8, "allgather") == "1"
    # 4 x 2 grid: rank = ri * 2 + ci; the peers of a rank hold the same feature slice, one per row block
    assert sdist.push_peers(5, 8, 2) == [1, 3, 5, 7]
    assert sdist.grid_coords(5, 8, 2) == (2, 1) and sdist.feature_slice(100, 2, 1) == (50, 100)
    rows_per, starts = sdist.row_partition(2449029, 4)
    assert rows_per == 612258 and starts.tolist() == [0, 612258, 1224516, 1836774, 2449029]
