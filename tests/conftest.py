import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_cls():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "classification.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_curves():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "curves.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_det():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "detection.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_reg():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "regression.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_binned():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "binned.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_multilabel():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "multilabel.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_consumers():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "consumers.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_fuzz():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "fuzz.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_atfixed():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "atfixed.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_logauc():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "logauc.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_coco_format():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "coco_format.npz"), allow_pickle=False)
