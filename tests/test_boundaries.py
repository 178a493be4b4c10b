"""Structural rules of the repo: the product never touches the oracle or the reference tree, and the GPU
paths cannot silently fall back to the CPU."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(top):
    for d, _, files in os.walk(os.path.join(ROOT, top)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(d, f)


def test_product_never_imports_or_loads_the_oracle():
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|.*libzs_oracle|.*zso_)", re.M)
    for path in _py_files("libzombsole_b200"):
        text = open(path, encoding="utf-8").read()
        code = "\n".join(l for l in text.split("\n") if not l.lstrip().startswith(("#", "//", "*")))
        # docstrings may MENTION the oracle; imports / symbol uses may not appear
        assert not re.search(r"^\s*(from\s+oracle\b|import\s+oracle\b)", code, re.M), path
        assert "libzs_oracle" not in code, path
        assert not re.search(r"\bzso_[a-z_]+\s*\(", code), path


def test_gpu_side_never_reads_the_reference_tree():
    # /root/reference does not exist on the GPU box: nothing the `-m gpu` tests, smoke() or bench.py execute may name it
    gpu_side = [os.path.join(ROOT, p) for p in (
        "bench.py", "__graft_entry__.py", "oracle/oracle.py", "tests/cuda_engine.py", "tests/parity_util.py",
        "tests/conftest.py", "tests/test_cuda_parity.py", "tests/test_cuda_properties.py", "tests/test_reference_api.py")]
    gpu_side += list(_py_files("libzombsole_b200"))
    for path in gpu_side:
        assert "/root/reference" not in open(path, encoding="utf-8").read(), path


def test_no_compat_layers_in_the_product():
    for path in _py_files("libzombsole_b200"):
        text = open(path, encoding="utf-8").read()
        assert "import triton" not in text and "torch.compile" not in text, path
