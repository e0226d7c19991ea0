"""Reference arm of bench.py: place the UNMODIFIED reference files of the propagation path under baseline/_ref/.

baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but travels to the GPU box
with the gpurun snapshot, where /root/reference does not exist.  The reference has no setup.py / pyproject, so there
is nothing to pip-install: the files SURVEY.md 8c lists are copied byte for byte, its prebuilt OpenMP library
(operators/csrc/libmatmul.so) included; if that binary cannot be loaded on the box, matmul.c is compiled in place
with the one gcc line of the survey.  Run by __graft_entry__.build() in the build container.
"""
import os
import shutil
import subprocess

REF = "/root/reference/Scalable Spectral Robust GNN"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "operators/base_operator.py",
    "operators/utils.py",
    "operators/graph_operator/symmetrical_simgraph_laplacian_operator.py",
    "operators/graph_operator/symmetrical_simgraph_ppr_operator.py",
    "operators/csrc/matmul.c",
    "operators/csrc/matmul.h",
    "operators/csrc/libmatmul.so",
]


def vendor() -> bool:
    """Copy the files; returns False (and does nothing) when /root/reference is absent (GPU box)."""
    if not os.path.isdir(REF):
        return os.path.isdir(DST)
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    return True


def ensure_libmatmul() -> str:
    """The reference loads ./csrc/libmatmul.so next to operators/utils.py.  Keep its prebuilt binary when it loads;
    otherwise build matmul.c in place (gcc -O3 -fopenmp -mavx2 -mfma)."""
    import ctypes
    so = os.path.join(DST, "operators/csrc/libmatmul.so")
    try:
        ctypes.CDLL(so)
        return "prebuilt"
    except OSError:
        src = os.path.join(DST, "operators/csrc/matmul.c")
        subprocess.run(["gcc", "-O3", "-fopenmp", "-mavx2", "-mfma", "-shared", "-fPIC", src, "-o", so], check=True)
        ctypes.CDLL(so)
        return "rebuilt from matmul.c"


if __name__ == "__main__":
    print("vendored" if vendor() else "reference not available", ensure_libmatmul() if os.path.isdir(DST) else "")
