"""Bandwidth kernels (LayerNorm, patchify, pooling, column sums, CE, AdamW) vs torch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

from conftest import elem_err, rel_err

pytestmark = pytest.mark.gpu


# rows >= 4096 take the shared-memory-ring backward (bulk async copies), fewer rows the register-resident one
@pytest.mark.parametrize("rows,dim", [(1576, 192), (1000, 384), (4097, 768), (577, 1024), (3, 768),
                                      (5000, 192), (4100, 384), (4096, 512), (4099, 1024), (50432, 768)])
def test_layernorm_fwd_bwd(cuda_device, rows, dim):
    from vision_transformers_torch_xla_b200 import _lib as L
    torch.manual_seed(0)
    x = torch.randn(rows, dim, device=cuda_device) * 2 + 0.5
    gamma = torch.randn(dim, device=cuda_device)
    beta = torch.randn(dim, device=cuda_device)
    y = torch.empty(rows, dim, device=cuda_device, dtype=torch.bfloat16)
    mean = torch.empty(rows, device=cuda_device)
    rstd = torch.empty(rows, device=cuda_device)
    L.layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, dim, 1e-6)
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (dim,), gr, br, 1e-6)
    assert elem_err(y.float(), ref) < 6e-3
    assert rel_err(mean, x.mean(-1)) < 1e-5
    assert rel_err(rstd, (x.var(-1, unbiased=False) + 1e-6).rsqrt()) < 1e-5

    dy = torch.randn(rows, dim, device=cuda_device).bfloat16()
    g_in = torch.randn(rows, dim, device=cuda_device)
    group = 7
    rs = torch.rand((rows + group - 1) // group, device=cuda_device) + 0.5
    g_out = torch.empty(rows, dim, device=cuda_device)
    gb = torch.empty(rows, dim, device=cuda_device, dtype=torch.bfloat16)
    dgamma = torch.zeros(dim, device=cuda_device)
    dbeta = torch.zeros(dim, device=cuda_device)
    L.layernorm_bwd(dy, x, mean, rstd, gamma, g_in, g_out, gb, rs, group, dgamma, dbeta, rows, dim)
    ref.backward(dy.float())
    want = g_in + xr.grad
    assert rel_err(g_out, want) < 1e-4
    assert elem_err(gb.float(), want * rs.repeat_interleave(group)[:rows, None]) < 6e-3
    assert rel_err(dgamma, gr.grad) < 1e-4
    assert rel_err(dbeta, br.grad) < 1e-4
    # without g_in / gb / rowscale
    g2 = torch.empty(rows, dim, device=cuda_device)
    L.layernorm_bwd(dy, x, mean, rstd, gamma, None, g2, None, None, 1, None, None, rows, dim)
    assert rel_err(g2, xr.grad) < 1e-4
    # in place on the residual gradient (g_out is g_in), the way the block backward calls it
    g3 = g_in.clone()
    L.layernorm_bwd(dy, x, mean, rstd, gamma, g3, g3, None, None, 1, None, None, rows, dim)
    assert torch.equal(g3, g_out)


def test_layernorm_strided_rows(cuda_device):
    """'token' pooling normalises only the cls rows: row pitch N*D."""
    from vision_transformers_torch_xla_b200 import _lib as L
    B, N, D = 5, 197, 192
    x = torch.randn(B, N, D, device=cuda_device)
    gamma = torch.randn(D, device=cuda_device)
    beta = torch.randn(D, device=cuda_device)
    y = torch.empty(B, D, device=cuda_device, dtype=torch.bfloat16)
    mean = torch.empty(B, device=cuda_device)
    rstd = torch.empty(B, device=cuda_device)
    L.layernorm_fwd(x, gamma, beta, y, mean, rstd, B, D, 1e-6, ld_x=N * D)
    assert elem_err(y.float(), F.layer_norm(x[:, 0], (D,), gamma, beta, 1e-6)) < 6e-3


@pytest.mark.parametrize("B,S", [(2, 224), (1, 384)])
def test_patchify(cuda_device, B, S):
    from vision_transformers_torch_xla_b200 import _lib as L
    img = torch.randn(B, 3, S, S, device=cuda_device)
    P = (S // 16) ** 2
    out = torch.empty(B * P, 768, device=cuda_device, dtype=torch.bfloat16)
    L.patchify(img, out, 16)
    ref = F.unfold(img, kernel_size=16, stride=16).transpose(1, 2).reshape(B * P, 768)  # (c, ph, pw) order
    assert torch.equal(out, ref.bfloat16())


def test_prefix_pool_embed_bwd(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    B, N, D, prefix = 33, 198, 192, 2
    x = torch.randn(B, N, D, device=cuda_device)
    tok = torch.randn(prefix, D, device=cuda_device)
    pos = torch.randn(N, D, device=cuda_device)
    x2 = x.clone()
    L.prefix_rows(x2, tok, pos, B, N, D, prefix)
    assert torch.allclose(x2[:, :prefix], (tok + pos[:prefix]).expand(B, -1, -1))
    assert torch.equal(x2[:, prefix:], x[:, prefix:])

    pooled = torch.empty(B, D, device=cuda_device)
    L.pool_fwd(x, pooled, B, N, D, prefix, 0)
    assert rel_err(pooled, x[:, prefix:].mean(1)) < 1e-5
    L.pool_fwd(x, pooled, B, N, D, prefix, 1)
    assert torch.equal(pooled, x[:, 0])

    dp = torch.randn(B, D, device=cuda_device)
    g = torch.full((B, N, D), float("nan"), device=cuda_device)
    L.pool_bwd(dp, g, B, N, D, prefix, 0)
    ref = torch.zeros(B, N, D, device=cuda_device)
    ref[:, prefix:] = dp[:, None] / (N - prefix)
    assert rel_err(g, ref) < 1e-6
    L.pool_bwd(dp, g, B, N, D, prefix, 1)
    ref = torch.zeros(B, N, D, device=cuda_device)
    ref[:, 0] = dp
    assert torch.equal(g, ref)

    gg = torch.randn(B, N, D, device=cuda_device)
    gp = torch.empty(B * (N - prefix), D, device=cuda_device, dtype=torch.bfloat16)
    dpos = torch.zeros(N, D, device=cuda_device)
    # the two token gradients live in separate rows of the flat gradient buffer and are ACCUMULATED into
    dcls, ddist = torch.ones(D, device=cuda_device), torch.full((D,), 2.0, device=cuda_device)
    L.embed_bwd(gg, gp, dpos, dcls, ddist, B, N, D, prefix)
    assert torch.equal(gp.view(B, N - prefix, D), gg[:, prefix:].bfloat16())
    assert rel_err(dpos, gg.sum(0)) < 1e-5
    assert rel_err(dcls - 1.0, gg[:, 0].sum(0)) < 1e-5 and rel_err(ddist - 2.0, gg[:, 1].sum(0)) < 1e-5
    dpos.zero_()
    L.embed_bwd(gg, gp, dpos, None, None, B, N, D, prefix)   # frozen tokens
    assert rel_err(dpos, gg.sum(0)) < 1e-5


@pytest.mark.parametrize("rows,cols", [(1576, 576), (5000, 3072), (100, 1000), (7, 768)])
def test_colsum(cuda_device, rows, cols):
    from vision_transformers_torch_xla_b200 import _lib as L
    x = torch.randn(rows, cols, device=cuda_device).bfloat16()
    out = torch.ones(cols, device=cuda_device)
    L.colsum_bf16(x, out, rows, cols)
    assert rel_err(out, 1 + x.float().sum(0)) < 1e-4


def test_ce_soft_targets(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    B, C = 64, 1000
    torch.manual_seed(0)
    logits = torch.randn(B, C, device=cuda_device) * 3
    soft = torch.softmax(torch.randn(B, C, device=cuda_device) * 2, -1)
    loss = torch.empty(1, device=cuda_device)
    dl = torch.empty(B, C, device=cuda_device)
    scratch = torch.empty(B, device=cuda_device)
    L.ce_fwd_bwd(logits, soft, None, 0.0, None, 0.0, 1.0, loss, dl, scratch)
    lr = logits.clone().requires_grad_(True)
    ref = torch.sum(-soft * F.log_softmax(lr, dim=-1), dim=-1).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    assert rel_err(dl, lr.grad) < 1e-4


@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_ce_hard_labels(cuda_device, smoothing):
    from vision_transformers_torch_xla_b200 import _lib as L
    B, C = 32, 1000
    logits = torch.randn(B, C, device=cuda_device) * 3
    labels = torch.randint(0, C, (B,), device=cuda_device)
    loss = torch.empty(1, device=cuda_device)
    dl = torch.empty(B, C, device=cuda_device)
    scratch = torch.empty(B, device=cuda_device)
    L.ce_fwd_bwd(logits, None, labels, smoothing, None, 0.0, 1.0, loss, dl, scratch)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, label_smoothing=smoothing)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    assert rel_err(dl, lr.grad) < 1e-4


def test_ce_distillation(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    B, C, alpha, T = 16, 1000, 0.7, 4.0
    logits = torch.randn(B, C, device=cuda_device) * 3
    teacher = torch.randn(B, C, device=cuda_device) * 3
    labels = torch.randint(0, C, (B,), device=cuda_device)
    loss = torch.empty(1, device=cuda_device)
    dl = torch.empty(B, C, device=cuda_device)
    scratch = torch.empty(B, device=cuda_device)
    L.ce_fwd_bwd(logits, None, labels, 0.1, teacher, alpha, T, loss, dl, scratch)
    lr = logits.clone().requires_grad_(True)
    base = F.cross_entropy(lr, labels, label_smoothing=0.1)
    kd = F.kl_div(F.log_softmax(lr / T, dim=1), F.softmax(teacher / T, dim=1), reduction="batchmean")
    ref = (1 - alpha) * base + alpha * T * T * kd
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    assert rel_err(dl, lr.grad) < 1e-4


def test_scale_cast(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    x = torch.randn(1003, device=cuda_device)[:1000]
    x = x.contiguous()
    s = torch.tensor([0.25], device=cuda_device)
    out = torch.empty(1000, device=cuda_device, dtype=torch.bfloat16)
    L.scale_cast_bf16(x, s, out)
    assert torch.equal(out, (x * 0.25).bfloat16())
    L.cast_bf16(x, out)
    assert torch.equal(out, x.bfloat16())


def test_adamw_matches_torch(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    torch.manual_seed(0)
    n, chunk = 64 * 37, 64
    p0 = torch.randn(n, device=cuda_device)
    groups = (torch.arange(n // chunk, device=cuda_device) % 2).to(torch.uint8)
    mask0 = groups.repeat_interleave(chunk) == 0
    pa = torch.nn.Parameter(p0[mask0].clone())
    pb = torch.nn.Parameter(p0[~mask0].clone())
    opt = torch.optim.AdamW([{"params": [pa], "weight_decay": 0.05, "lr": 1e-2},
                             {"params": [pb], "weight_decay": 0.0, "lr": 5e-3}], betas=(0.9, 0.999), eps=1e-8)
    p = p0.clone()
    m = torch.zeros(n, device=cuda_device)
    v = torch.zeros(n, device=cuda_device)
    shadow = torch.empty(n, device=cuda_device, dtype=torch.bfloat16)
    ema = p0.clone()
    ema_ref = p0.clone()
    for step in range(1, 6):
        g = torch.randn(n, device=cuda_device)
        pa.grad = g[mask0].clone()
        pb.grad = g[~mask0].clone()
        opt.step()
        gg = g.clone() * 4.0
        L.adamw_flat(p, gg, m, v, shadow, ema, groups, chunk, [1e-2, 5e-3], [0.05, 0.0], 0.9, 0.999, 1e-8, step,
                     grad_scale=0.25, ema_decay=0.99, zero_grad=True)
        assert float(gg.abs().max()) == 0.0
        want = torch.empty(n, device=cuda_device)
        want[mask0] = pa.detach()
        want[~mask0] = pb.detach()
        ema_ref = 0.99 * ema_ref + 0.01 * want
        assert rel_err(p, want) < 1e-5
        assert torch.equal(shadow, p.bfloat16())
        assert rel_err(ema, ema_ref) < 1e-5


def test_sumsq(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    x = torch.randn(1_000_003, device=cuda_device)[:1_000_000].contiguous()
    out = torch.zeros(1, device=cuda_device)
    L.sumsq(x, out)
    assert abs(out.item() - x.double().pow(2).sum().item()) < 1e-3 * x.numel() ** 0.5 * 10
    xb = torch.randn(1_000_008, device=cuda_device).bfloat16()[:1_000_003].clone()
    out.zero_()
    L.sumsq(xb, out)
    assert abs(out.item() - xb.double().pow(2).sum().item()) < 1e-3 * xb.numel() ** 0.5 * 10
    # deterministic: no floating-point atomics (data-parallel ranks derive their clip coefficient from this)
    big = torch.randn(86_567_656, device=cuda_device)
    results = []
    for _ in range(5):
        o = torch.zeros(1, device=cuda_device)
        L.sumsq(big, o)
        results.append(o.item())
    assert len(set(results)) == 1, results


def test_adamw_bf16_gradient_source_and_device_scale(cuda_device):
    """vitk_adamw_flat reading the gradient from the bf16 copy the data-parallel layer all-reduces, times a device-side
    scalar (the clip coefficient): equals torch.optim.AdamW fed bf16(g) * grad_scale * coef; the fp32 buffer is zeroed."""
    from vision_transformers_torch_xla_b200 import _lib as L
    torch.manual_seed(0)
    n = 64 * 21
    p0 = torch.randn(n, device=cuda_device)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pr], lr=1e-2, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device=cuda_device), torch.zeros(n, device=cuda_device)
    coef = torch.tensor([0.37], device=cuda_device)
    for step in range(1, 4):
        g = torch.randn(n, device=cuda_device)
        g16 = (g * 8.0).bfloat16()
        pr.grad = g16.float() * 0.125 * 0.37
        opt.step()
        junk = torch.full((n,), 123.0, device=cuda_device)   # the fp32 buffer must not be read, only zeroed
        L.adamw_flat(p, junk, m, v, None, None, None, 64, [1e-2], [0.05], 0.9, 0.999, 1e-8, step, grad_scale=0.125,
                     zero_grad=True, g_bf16=g16, grad_scale_dev=coef)
        assert float(junk.abs().max()) == 0.0
        assert rel_err(p, pr.detach()) < 1e-5


def test_scale_f32_and_clip_coef(cuda_device):
    from vision_transformers_torch_xla_b200 import _lib as L
    x = torch.randn(1003, device=cuda_device)
    want = x * 0.25
    L.scale_f32_(x, torch.tensor(0.25, device=cuda_device))
    assert torch.equal(x, want)
    buf = torch.tensor([16.0, 0.0, 0.0], device=cuda_device)          # sum of squares 16 -> norm of the SUM 4
    L.clip_coef(buf[0:1], 0.5, 1.0, buf[1:2], buf[2:3])               # mean over 2 replicas: norm 2 -> coef 1 / (2 + 1e-6)
    assert abs(buf[2].item() - 2.0) < 1e-6 and abs(buf[1].item() - 1.0 / (2.0 + 1e-6)) < 1e-6
    L.clip_coef(buf[0:1], 0.5, 10.0, buf[1:2], None)
    assert buf[1].item() == 1.0
