import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the oracle runs next to the kernels as the fp32 reference: keep TF32 out of its matmuls and convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def report(name, a, b):
    """One line with every metric the tolerances are stated in, so that the margin is visible in the test log."""
    line = (f"[parity] {name}: max/rms {rel_err(a, b):.3e}  elem {elem_err(a, b):.3e}  rms {rms_err(a, b):.3e}  "
            f"cos {cos_sim(a, b):.6f}")
    print(line)
    return line


def rel_err(a, b):
    """max |a-b| normalised by the RMS of the reference (north_star's per-tensor metric)."""
    import torch

    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    rms = b.pow(2).mean().sqrt().clamp_min(1e-30)
    return float((a - b).abs().max() / rms)


def elem_err(a, b):
    """max |a-b| / (|b| + rms(b)): elementwise-relative error with an RMS floor.

    One bf16 rounding of the output costs at most 2^-8 = 3.9e-3 here regardless of how far into the
    tail of the distribution the element sits, so bf16-output kernels are checked with this metric
    (a stricter statement than the north-star's max/RMS <= 2e-2, which is asserted as well).
    """
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    rms = b.pow(2).mean().sqrt().clamp_min(1e-30)
    return float(((a - b).abs() / (b.abs() + rms)).max())


def rms_err(a, b):
    """rms(a-b) / rms(b): the typical (not worst-case) relative error of a tensor."""
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))


def cos_sim(a, b):
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
