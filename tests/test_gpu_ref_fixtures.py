"""GPU: the CUDA path (through the C ABI) against OUTPUTS OF THE REFERENCE'S OWN CODE (tests/golden/ref_fixtures.pt,
written by tests/golden/make_ref_fixtures.py from the reference sources; see tests/test_ref_fixtures.py for the CPU half).

Tolerances (north_star: bf16 kernels vs the fp32 reference, per-tensor max / RMS <= 2e-2): activations and logits
max/rms <= 2e-2; gradients rms <= 1.5e-2 and cosine >= 0.999 with the worst element bounded at max/rms <= 1.5e-1 (a
weight gradient of the micro model is a sum over only 20 token rows: one bf16 rounding of a large term is not averaged
away); every measured value is printed.  DropPath masks are REPLAYED: the fixture
holds the masks the reference run drew, ops.mask_source injects them, so drop_path > 0 is compared exactly.
"""
import os
import sys

import pytest
import torch

from conftest import cos_sim, rel_err, report, rms_err

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_inputs as RI  # noqa: E402

pytestmark = pytest.mark.gpu
FX_PATH = os.path.join(HERE, "golden", "ref_fixtures.pt")


@pytest.fixture(scope="module")
def fx():
    return torch.load(FX_PATH, weights_only=False)


class replay_masks:
    """ops.mask_source hook: hand out the DropPath masks of the reference run, in the reference's draw order."""

    def __init__(self, masks, dev):
        self.masks, self.dev = [m.to(dev) for m in masks], dev

    def __enter__(self):
        from vision_transformers_torch_xla_b200 import ops

        def source(drop_probs, B, device):
            rows, it = [], iter(self.masks)
            for p in drop_probs:
                rows.append(next(it) if p > 0.0 else torch.ones(B, device=device))
            assert next(it, None) is None, "the reference drew more masks than the model has DropPath layers"
            return torch.stack(rows).float()

        self.ops = ops
        ops.mask_source = source
        return self

    def __exit__(self, *exc):
        self.ops.mask_source = None
        return False


def state_for(case, name, fx):
    """The scenario's weights: stored whole for 'avg' (and shared by 'token'); the others are rebuilt with the oracle
    from the same seeds and verified against the reference's checksums."""
    from oracle import vit_oracle as O

    if name in ("avg", "token"):
        sd = dict(fx["micro"]["avg"]["state_dict"])
        if name == "token":
            sd = {k.replace("fc_norm.", "norm."): v for k, v in sd.items()}
        return sd
    cls = O.VisionTransformerDistilled if name == "distilled" else O.VisionTransformer
    torch.manual_seed(case["seeds"]["init"])
    m = cls(**case["kwargs"])
    RI.perturb(m, case["seeds"]["perturb"])
    sd = m.state_dict()
    got = RI.checksums(list(sd.items()))
    assert torch.allclose(got, case["state_checksums"], rtol=1e-9, atol=1e-12)
    return sd


def run_mine(model, x, loss_fn):
    acts = []
    hooks = [b.register_forward_hook(lambda m, i, o: acts.append(o.detach().clone())) for b in model.blocks]
    out = model(x)
    loss = loss_fn(out)
    loss.backward()
    for h in hooks:
        h.remove()
    return out, loss, acts


def check_grads(model, want, tag):
    worst = (0.0, "")
    for n, p in model.named_parameters():
        g, w = p.grad, want[n].to(p.device)
        r, c, e = rms_err(g, w), cos_sim(g, w), rel_err(g, w)
        worst = max(worst, (e, n))
        assert r < 1.5e-2 and c > 0.999, (tag, n, r, c)
        assert e < 1.5e-1, (tag, n, e)
    print(f"[parity] {tag}: worst parameter-gradient max/rms {worst[0]:.3e} ({worst[1]})")


@pytest.mark.parametrize("name", ["avg", "token", "avg_ls_dp"])
def test_cuda_path_matches_reference_vit_micro(cuda_device, fx, name):
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import VisionTransformer

    case = fx["micro"][name]
    model = VisionTransformer(**case["kwargs"]).to(cuda_device)
    model.load_state_dict(state_for(case, name, fx))
    model.train()
    x, tgt = case["x"].to(cuda_device), case["target"].to(cuda_device)
    with replay_masks(case["masks"], cuda_device):
        out, loss, acts = run_mine(model, x, lambda o: SoftTargetCrossEntropy()(o, tgt))
    report(f"micro/{name} logits", out, case["logits"].to(cuda_device))
    assert rel_err(out, case["logits"].to(cuda_device)) < 2e-2
    assert abs(loss.item() - case["loss"].item()) < 2e-3 * abs(case["loss"].item())
    assert len(acts) == len(case["acts"])
    for i, (a, b) in enumerate(zip(acts, case["acts"])):
        assert rel_err(a, b.to(cuda_device)) < 2e-2, f"activation of block {i}: {rel_err(a, b.to(cuda_device))}"
    check_grads(model, case["grads"], f"micro/{name}")
    model.eval()
    with torch.no_grad():
        assert rel_err(model(x), case["logits_eval"].to(cuda_device)) < 2e-2


def test_cuda_path_matches_reference_distilled_micro(cuda_device, fx):
    from vision_transformers_torch_xla_b200.losses import DistillationLoss, LabelSmoothingCrossEntropy
    from vision_transformers_torch_xla_b200.models.deit import VisionTransformerDistilled

    case = fx["micro"]["distilled"]
    model = VisionTransformerDistilled(**case["kwargs"]).to(cuda_device)
    model.load_state_dict(state_for(case, "distilled", fx))
    model.train()
    model.set_distilled_training(True)
    x = case["x"].to(cuda_device)
    labels, teacher = case["labels"].to(cuda_device), case["teacher"].to(cuda_device)
    # 0.5 CE(cls, y) + 0.5 CE(dist, argmax teacher): DeiT's hard distillation == DistillationLoss(hard, alpha 0.5, no smoothing)
    crit = DistillationLoss(LabelSmoothingCrossEntropy(0.0), alpha=0.5, temperature=1.0, hard=True)
    out, loss, acts = run_mine(model, x, lambda o: crit((o, teacher), labels))
    assert isinstance(out, tuple)
    for j in range(2):
        assert rel_err(out[j], case["logits"][j].to(cuda_device)) < 2e-2
    assert abs(loss.item() - case["loss"].item()) < 2e-3 * abs(case["loss"].item())
    for a, b in zip(acts, case["acts"]):
        assert rel_err(a, b.to(cuda_device)) < 2e-2
    check_grads(model, case["grads"], "micro/distilled")
    model.set_distilled_training(False)
    assert rel_err(model(x), case["logits_train_avg"].to(cuda_device)) < 2e-2
    model.eval()
    with torch.no_grad():
        assert rel_err(model(x), case["logits_eval"].to(cuda_device)) < 2e-2


@pytest.mark.parametrize("name", ["vit_tiny_patch16_224", "deit_tiny_distilled_patch16_224", "my_vit_mini", "my_vit_xs"])
def test_named_configs_with_drop_path_match_reference(cuda_device, fx, name):
    """The reference's own entrypoints with drop_path_rate 0.1 (the launch value, run_train.sh:58): config 1 (ViT-Ti),
    distilled DeiT-Ti, my_vit_mini (head_dim 48, my_vit.py:85-95) and my_vit_xs (head_dim 72, my_vit.py:97-106) on a seeded 2-image batch with the reference's
    masks replayed: logits, loss, the last block's activations, every small gradient elementwise, all gradients by
    their energy."""
    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import create_model

    rec = fx["named"][name]
    x, tgt = (t.to(cuda_device) for t in RI.named_inputs())
    torch.manual_seed(42)
    ref = O.create_model(name, **rec["kwargs"])   # same seeded init as the reference (pinned by the CPU tests)
    assert torch.allclose(RI.checksums(list(ref.state_dict().items())), rec["init_checksums"], rtol=1e-9, atol=1e-12)
    model = create_model(name, **rec["kwargs"]).to(cuda_device)
    model.load_state_dict(ref.state_dict())
    model.train()
    ce = SoftTargetCrossEntropy()
    if name.startswith("deit_"):
        model.set_distilled_training(True)
        fn = lambda o: ce(o[0], tgt) + ce(o[1], tgt.flip(0))  # noqa: E731
    else:
        fn = lambda o: ce(o, tgt)  # noqa: E731
    with replay_masks(rec["masks"], cuda_device):
        out, loss, acts = run_mine(model, x, fn)
    outs = out if isinstance(out, tuple) else (out,)
    wants = rec["logits"] if isinstance(rec["logits"], tuple) else (rec["logits"],)
    for a, b in zip(outs, wants):
        report(f"{name} logits", a, b.to(cuda_device))
        # 'avg' pooling averages 196 tokens' errors away; the distilled model's heads read ONE token each: 2x the error
        assert rel_err(a, b.to(cuda_device)) < (3e-2 if name.startswith("deit_") else 2e-2)
        assert rms_err(a, b.to(cuda_device)) < 1e-2
    assert abs(loss.item() - rec["loss"].item()) < 2e-3 * abs(rec["loss"].item())
    assert rel_err(acts[-1][:, :3, :16], rec["act_last_slice"].to(cuda_device)) < 2e-2
    grads = {n: p.grad for n, p in model.named_parameters()}
    for k, v in rec["grads_small"].items():
        if k.endswith("attn.qkv.bias") or v.abs().max() == 0:
            continue   # (the key-bias third of qkv.bias has an exactly-zero true gradient: pure rounding noise)
        assert rms_err(grads[k], v.to(cuda_device)) < 2e-2 and cos_sim(grads[k], v.to(cuda_device)) > 0.999, k
    energy = RI.checksums(list(grads.items()))[:, 1]
    want = rec["grad_checksums"][:, 1]
    rel = ((energy - want).abs() / want.clamp_min(1e-30))
    keep = torch.tensor([not k.endswith("attn.qkv.bias") for k in rec["grad_keys"]])
    print(f"[parity] {name}: worst relative deviation of a gradient's sum of squares {float(rel[keep].max()):.3e}")
    assert float(rel[keep].max()) < 3e-2


@pytest.mark.parametrize("name", ["uf1", "uf2"])
def test_engine_matches_reference_train_one_epoch(cuda_device, fx, name):
    """engine.train_one_epoch on the CUDA path against the reference's own train_one_epoch (engine.py:19-333, eager
    branch) on the same model, batches, schedules and update_freq over two epochs: per-micro-batch losses, returned
    meters, the weight-decay quirk, and the weights after 10 / 6 AdamW steps."""
    from vision_transformers_torch_xla_b200 import engine, optim_factory
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import VisionTransformer

    rec = fx["engine"][name]
    batches, _ = RI.engine_inputs()
    batches = batches[:rec["n_micro"]]
    init = fx["micro"]["avg"]["state_dict"]
    model = VisionTransformer(**rec["kwargs"]).to(cuda_device)
    model.load_state_dict(init)

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas = "adamw", rec["args"]["lr"], rec["args"]["weight_decay"], 1e-8, None

    opt = optim_factory.create_optimizer(Args, model)
    losses = []

    class Crit(SoftTargetCrossEntropy):
        def forward(self, o, t):
            loss = super().forward(o, t)
            losses.append(loss.detach())
            return loss

    lr_s, wd_s = rec["lr_schedule"].numpy(), rec["wd_schedule"].numpy()
    half, spe = rec["n_micro"] // 2, rec["steps_per_epoch"]
    for epoch in range(2):
        st = engine.train_one_epoch(model, Crit(), batches[epoch * half:(epoch + 1) * half], opt, cuda_device, epoch, None,
                                    mixup_fn=lambda s_, t_: (s_, t_), start_steps=epoch * spe, lr_schedule_values=lr_s,
                                    wd_schedule_values=wd_s, num_training_steps_per_epoch=spe,
                                    update_freq=rec["update_freq"], log_freq=1)
        want = rec["stats"][epoch]
        assert set(st) == set(want), (st, want)
        assert abs(st["loss"] - want["loss"]) < 5e-3 * abs(want["loss"]) and abs(st["lr"] - want["lr"]) < 1e-12
    got = torch.stack(losses).float().cpu()
    rel = ((got - rec["losses"]).abs() / rec["losses"].abs()).max().item()
    print(f"[parity] engine/{name}: max relative loss deviation over {len(losses)} micro-batches {rel:.3e}")
    assert rel < 5e-3
    assert [g["weight_decay"] for g in opt.param_groups] == rec["final_group_wd"]
    assert [g["lr"] for g in opt.param_groups] == rec["final_group_lr"]
    sd = model.state_dict()
    worst = 0.0
    for k, v in rec["final_small"].items():
        if k.endswith("attn.qkv.bias"):
            continue
        v, v0 = v.to(cuda_device), init[k].to(cuda_device)
        # agreement of the UPDATE (final - initial): Adam's normalised steps make this the sensitive quantity
        upd = rms_err(sd[k] - v0, v - v0)
        worst = max(worst, upd)
        assert upd < 0.15 and rms_err(sd[k], v) < 2e-2, (k, upd)
    print(f"[parity] engine/{name}: worst rms deviation of a parameter's total update {worst:.3e}")


def test_evaluate_matches_reference(cuda_device, fx):
    """engine.evaluate against the reference's evaluate() (engine.py:339-430) on the weights its uf1 run ended with is
    covered on the CPU for the oracle; here the CUDA path evaluates the INITIAL micro weights against the oracle."""
    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200 import engine
    from vision_transformers_torch_xla_b200.models import VisionTransformer

    _, hard = RI.engine_inputs()
    kw = fx["engine"]["uf1"]["kwargs"]
    sd = fx["micro"]["avg"]["state_dict"]
    ref = O.VisionTransformer(**kw)
    ref.load_state_dict(sd)
    model = VisionTransformer(**kw).to(cuda_device)
    model.load_state_dict(sd)
    want = O.evaluate(hard, ref)
    got = engine.evaluate(hard, model, cuda_device)
    assert set(got) == set(want) == {"loss", "acc1", "acc5"}
    assert abs(got["loss"] - want["loss"]) < 5e-3 * abs(want["loss"])
    assert abs(got["acc1"] - want["acc1"]) <= 100.0 / 12 + 1e-6 and abs(got["acc5"] - want["acc5"]) <= 100.0 / 12 + 1e-6


def test_distillation_loss_matches_reference(cuda_device, fx):
    """The fused KD cross-entropy kernel against values and gradients the reference's closure-local DistillationLoss
    (main.py:939-968) produced."""
    from vision_transformers_torch_xla_b200.losses import CrossEntropyLoss, DistillationLoss, SoftTargetCrossEntropy

    kd = fx["host"]["kd"]
    s, t, y, ysoft = (v.to(cuda_device) for v in RI.kd_inputs())
    for (a, T), want in kd["cases"].items():
        s1 = s.clone().requires_grad_(True)
        l1 = DistillationLoss(CrossEntropyLoss(), a, T)((s1, t), y)
        l1.backward()
        assert abs(l1.item() - want["hard_labels"][0].item()) < 1e-5 * max(1.0, abs(want["hard_labels"][0].item()))
        assert rel_err(s1.grad, want["hard_labels"][1].to(cuda_device)) < 1e-4
        s2 = s.clone().requires_grad_(True)
        l2 = DistillationLoss(SoftTargetCrossEntropy(), a, T)((s2, t), ysoft)
        l2.backward()
        assert abs(l2.item() - want["soft_labels"][0].item()) < 1e-5 * max(1.0, abs(want["soft_labels"][0].item()))
        assert rel_err(s2.grad, want["soft_labels"][1].to(cuda_device)) < 1e-4
        l3 = DistillationLoss(CrossEntropyLoss(), a, T)(s, y)
        assert abs(l3.item() - want["tensor_input"].item()) < 1e-5 * max(1.0, abs(want["tensor_input"].item()))
