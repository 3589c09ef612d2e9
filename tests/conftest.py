import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def from_bits(a: np.ndarray) -> torch.Tensor:
    """int16 bit patterns -> bf16 tensor"""
    return torch.from_numpy(np.ascontiguousarray(a)).view(torch.bfloat16)


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


GRAD_SAMPLES = 2048


def grad_sample_index(numel: int) -> np.ndarray:
    """Indices of the gradient entries the titok_grads_* fixtures keep per parameter (tests/golden/make_golden.py)."""
    stride = max(1, -(-numel // GRAD_SAMPLES))
    return np.arange(0, numel, stride)


def param_checksum(sd) -> np.ndarray:
    return np.array([float(v.double().sum()) for _, v in sorted(sd.items())] +
                    [float(v.double().abs().sum()) for _, v in sorted(sd.items())])


def checksum_matches(sd, want: np.ndarray) -> bool:
    """float64 sums are order-dependent in the last bits across CPUs (vector width): compare to 1e-12 relative."""
    return bool(np.allclose(param_checksum(sd), want, rtol=1e-12, atol=1e-12))


def build_model(stress: bool, levels=(7, 5, 5, 5, 5), patch=(4, 8, 8), enc="tiny", dec="tiny", seed=42):
    """titok_video_b200.TiTok with the weights the golden fixtures were generated with (CPU, fp32)."""
    import titok_video_b200 as T
    from titok_video_b200.config import tiny_config
    from oracle import titok_oracle as O

    torch.manual_seed(seed)
    m = T.TiTok(tiny_config(levels, patch, enc, dec))
    if stress:
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        O.stress_init_(sd, 1)
        m.load_state_dict(sd)
    return m


@pytest.fixture(scope="session")
def golden_default():
    return load_golden("titok_default")


@pytest.fixture(scope="session")
def golden_stress():
    return load_golden("titok_stress")
