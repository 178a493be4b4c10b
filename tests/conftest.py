import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> (meta dict, [trace dict per env])"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    seed, base, E, T, mes = (int(v) for v in z["meta"])
    traces = []
    for e in range(E):
        pre = "e%d_" % e
        traces.append({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
    return dict(seed=seed, base=base, E=E, T=T, max_episode_steps=mes), traces


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


@pytest.fixture(scope="session")
def have_cuda():
    import torch
    return torch.cuda.is_available()
