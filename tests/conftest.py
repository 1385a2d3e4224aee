import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(REPO, "tests", "golden", "cldm_v15_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_vae():
    """Reference Decoder outputs (tests/golden/make_golden_vae.py)."""
    return dict(np.load(os.path.join(REPO, "tests", "golden", "vae_decoder_golden.npz")))


@pytest.fixture(scope="session")
def vae_state_dict_cpu():
    from prompt_diffusion_b200.synth import synthetic_vae_state_dict
    return synthetic_vae_state_dict(seed=0)


@pytest.fixture(scope="session")
def golden_clip():
    """transformers.CLIPTextModel outputs (tests/golden/make_golden_clip.py)."""
    return dict(np.load(os.path.join(REPO, "tests", "golden", "clip_text_golden.npz")))


@pytest.fixture(scope="session")
def clip_state_dict_cpu():
    from prompt_diffusion_b200.synth import synthetic_clip_state_dict
    return synthetic_clip_state_dict(seed=0)


@pytest.fixture(scope="session")
def cfg():
    from prompt_diffusion_b200.config import CLDM_V15
    return CLDM_V15


@pytest.fixture(scope="session")
def state_dict_cpu(cfg):
    """The procedural checkpoint (1.22 B params fp32, ~4.9 GB) — built once per session."""
    from prompt_diffusion_b200.synth import synthetic_state_dict
    return synthetic_state_dict(cfg, seed=0)


def rel_l2(a, b):
    import torch
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
