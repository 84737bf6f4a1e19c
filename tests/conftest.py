import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def rel_l2(a, b):
    import torch
    cv = lambda t: torch.view_as_real(t.detach().cpu().to(torch.complex128)) if t.is_complex() else t.detach().double().cpu()
    a, b = cv(a), cv(b)
    if a.shape != b.shape:      # one side complex, the other its (re, im) real view
        assert a.numel() == b.numel(), (a.shape, b.shape)
        b = b.reshape(a.shape)
    return (torch.linalg.norm(a - b) / torch.linalg.norm(b).clamp_min(1e-30)).item()
