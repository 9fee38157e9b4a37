"""Device Mixup / CutMix against the oracle's restatement of timm.data.Mixup (same NumPy seed -> same lam / box):
the mixed images and the soft targets must be bit-identical (fp32 elementwise work)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mixup_alpha,cutmix_alpha,prob,seed", [(0.8, 0.0, 1.0, 0), (0.0, 1.0, 1.0, 1), (0.8, 1.0, 1.0, 2),
                                                                 (0.8, 1.0, 1.0, 3), (0.8, 1.0, 0.5, 4), (0.8, 1.0, 0.5, 5)])
@pytest.mark.parametrize("shape", [(8, 3, 224, 224), (6, 3, 32, 36)])
def test_mixup_matches_oracle_bitwise(cuda_device, mixup_alpha, cutmix_alpha, prob, seed, shape):
    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200.mixup import Mixup

    kw = dict(mixup_alpha=mixup_alpha, cutmix_alpha=cutmix_alpha, prob=prob, switch_prob=0.5, label_smoothing=0.1, num_classes=1000)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g)
    y = torch.randint(0, 1000, (shape[0],), generator=g)
    np.random.seed(seed)
    xr, tr = O.Mixup(**kw)(x.clone().to(cuda_device), y.to(cuda_device))
    np.random.seed(seed)
    xm, tm = Mixup(**kw)(x.clone().to(cuda_device), y.to(cuda_device))
    assert torch.equal(xm, xr)
    assert torch.equal(tm, tr)
    assert abs(float(tm.sum(1).mean()) - 1.0) < 1e-5


def test_mixup_contract(cuda_device):
    from vision_transformers_torch_xla_b200._lib import VitkError
    from vision_transformers_torch_xla_b200.mixup import Mixup

    with pytest.raises(NotImplementedError):
        Mixup(mode="elem")
    with pytest.raises(NotImplementedError):
        Mixup(cutmix_minmax=(0.2, 0.8))
    with pytest.raises(VitkError):
        Mixup()(torch.zeros(2, 3, 8, 8), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(AssertionError):
        Mixup()(torch.zeros(3, 3, 8, 8, device=cuda_device), torch.zeros(3, dtype=torch.int64))
