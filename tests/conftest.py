import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def golden_names(prefix=None, exclude_prefix=("init_", "builder_", "reader_", "writer")):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    names = [n for n in names if not n.startswith(tuple(exclude_prefix))]
    if prefix is not None:
        names = [n for n in names if n.startswith(prefix)]
    return names


def load_golden(name):
    """-> dict(opt=..., params={state_dict key: array}, grads={...}, inputs..., outputs...)"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {"opt": {}, "params": {}, "grads": {}}
    for k in z.files:
        if k.startswith("opt_"):
            rec["opt"][k[4:]] = z[k].item()
        elif k.startswith("param."):
            rec["params"][k[6:]] = z[k]
        elif k.startswith("grad."):
            rec["grads"][k[5:]] = z[k]
        else:
            rec[k] = z[k]
    return rec


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
