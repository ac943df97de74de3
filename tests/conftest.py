import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: long-running CPU oracle checks (opt-in via SR_RUN_SLOW=1)")


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "slow" in item.keywords and os.environ.get("SR_RUN_SLOW", "0") != "1":
            item.add_marker(pytest.mark.skip(reason="set SR_RUN_SLOW=1"))


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_tiny():
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, "tiny.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle_tiny():
    """HF-eager fp32 oracle, reduced depth (2 ViT blocks / 2 decoder layers), seeded weights."""
    from oracle import hf_oracle as O

    return O.build_model("tiny")


@pytest.fixture(scope="session")
def tiny_batch():
    from oracle import hf_oracle as O

    return O.make_batch([0, 1], [224, 336], "tiny", n_text=8)


@pytest.fixture(scope="session")
def model_3b():
    """Full-depth 3B geometry on the GPU, weights from the seeded HF init (CPU build, ~2 min) -- shared by the 3B golden tests."""
    import torch
    from oracle import hf_oracle as O
    from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration as Model

    torch.set_num_threads(os.cpu_count())
    hf = O.build_model("3b")
    model = Model.from_hf_model(hf, device="cuda")
    del hf
    return model
