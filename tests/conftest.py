import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def testdata_frames():
    import numpy as np
    fr = np.load(GOLDEN / "frames_testdata.npz")
    K = fr["K"]
    return dict(bgr=fr["bgr"], depth=fr["depth"], gt=fr["gt"], K=tuple(float(v) for v in K),
                depth_scale=float(fr["depth_scale"]))
