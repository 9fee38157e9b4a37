"""Seeded inputs shared by tests/golden/make_ref_fixtures.py (which runs the reference on them, in the build
container) and the tests that replay the same scenarios on the oracle / the CUDA path.  torch's CPU generator is
deterministic for a given torch build; every regenerated tensor is verified against a checksum stored in the fixture."""
import torch

MICRO = dict(img_size=32, patch_size=16, embed_dim=64, depth=2, num_heads=1, num_classes=16)
SMALL = 4096   # tensors up to this many elements are stored whole in the fixture, larger ones as checksums


def perturb(model, seed):
    """non-trivial LayerNorm / bias / LayerScale values so that every parameter matters"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") or "norm" in n or n.endswith("gamma"):
                p.add_(0.1 * torch.randn(p.shape, generator=g).to(p.device))


def checksum(t):
    t = t.detach().double().flatten().cpu()
    return torch.stack([t.sum(), t.pow(2).sum(), t[0], t[t.numel() // 2], t[-1]])


def checksums(named):
    """[n_tensors, 5] float64 (sum, sum of squares, first / middle / last element), rows in the given key order"""
    return torch.stack([checksum(v) for _, v in named])


def micro_inputs():
    g = torch.Generator().manual_seed(99)
    x = torch.randn(4, 3, 32, 32, generator=g)
    tgt = torch.softmax(torch.randn(4, 16, generator=g) * 2, -1)
    labels = torch.randint(0, 16, (4,), generator=g)
    teacher = torch.randn(4, 16, generator=g)
    return x, tgt, labels, teacher


def named_inputs():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 224, 224, generator=g)
    tgt = torch.softmax(torch.randn(2, 1000, generator=g) * 3, -1)
    return x, tgt


def engine_inputs():
    g = torch.Generator().manual_seed(17)
    batches = [(torch.randn(4, 3, 32, 32, generator=g), torch.softmax(torch.randn(4, 16, generator=g) * 2, -1))
               for _ in range(12)]
    hard_batches = [(torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 16, (4,), generator=g)) for _ in range(3)]
    return batches, hard_batches


def kd_inputs():
    g = torch.Generator().manual_seed(3)
    s, t = torch.randn(6, 10, generator=g), torch.randn(6, 10, generator=g) * 2
    y = torch.randint(0, 10, (6,), generator=g)
    ysoft = torch.softmax(torch.randn(6, 10, generator=g), -1)
    return s, t, y, ysoft
