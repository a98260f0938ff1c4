import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def work_dir(tmp_path_factory):
    d = os.path.join(ROOT, "build", "tests")
    os.makedirs(d, exist_ok=True)
    return d


@pytest.fixture(scope="session")
def oracle_player():
    """The CPU restatement player (oracle/_build/oracle_player); built on demand with plain g++."""
    from softglrender_b200 import workloads
    if not os.path.exists(workloads.ORACLE_PLAYER):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restate"], check=True)
    return workloads.ORACLE_PLAYER


@pytest.fixture(scope="session")
def checker_player(oracle_player):
    """Strongest checker available: the compiled reference (deterministic single-worker build) when it travelled with
    the repo snapshot, else the CPU restatement."""
    from softglrender_b200 import workloads
    if os.path.exists(workloads.REF_PLAYER_ST):
        return workloads.REF_PLAYER_ST
    return oracle_player


def compare_outputs(ref, got, color_frac=0.999):
    """Parity bar of BASELINE.json: depth/coverage bit-exact, colour within 1/255 on >= 99.9 % of pixels."""
    report = {}
    for tag, x in ref.items():
        assert tag in got, "missing output " + tag
        y = got[tag]
        assert x.shape == y.shape, (tag, x.shape, y.shape)
        if x.dtype == np.uint8:
            d = np.abs(x.astype(np.int32) - y.astype(np.int32)).max(axis=-1)
            frac = float((d <= 1).mean())
            report[tag] = dict(within1=frac, exact=float((d == 0).mean()), max=int(d.max()))
            assert frac >= color_frac, "%s: only %.5f of pixels within 1/255 (max err %d)" % (tag, frac, d.max())
        else:
            eq = x.view(np.uint32) == y.view(np.uint32)
            report[tag] = dict(bit_exact=float(eq.mean()), mismatches=int((~eq).sum()))
            assert eq.all(), "%s: %d depth samples differ" % (tag, int((~eq).sum()))
    return report
