"""CPU: the oracle (and the product's host-side logic) against OUTPUTS OF THE REFERENCE'S OWN CODE.

tests/golden/ref_fixtures.pt was written by tests/golden/make_ref_fixtures.py, which ast-extracts the hot-path
definitions from /root/reference (models/vision_transformer.py, deit.py, my_vit.py, utils, optim_factory.py,
engine.py, main.py) and executes them with only the absent pip-timm leaf layers bound to the oracle's
restatements.  These tests are what turns "parity unpinned" into "pinned" for the in-tree half of the path:
the oracle must reproduce the reference's numbers at <= 1e-6 (same fp32 torch ops, so differences are only
summation order); the CUDA path is then held to the same fixtures in tests/test_gpu_ref_fixtures.py.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_inputs as RI  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

FX_PATH = os.path.join(HERE, "golden", "ref_fixtures.pt")
TOL = 1e-6


@pytest.fixture(scope="module")
def fx():
    return torch.load(FX_PATH, weights_only=False)


def close(a, b, tol=TOL):
    a, b = a.detach().double(), b.detach().double()
    scale = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / scale <= tol


def check_sums(named, want, tol=1e-9):
    got = RI.checksums(named)
    assert got.shape == want.shape
    scale = want.abs().clamp_min(1e-12)
    bad = ((got - want).abs() / scale > tol).any(dim=1).nonzero().flatten().tolist()
    assert not bad, ("checksum mismatch at rows", bad[:5], [named[i][0] for i in bad[:5]])


def oracle_micro(case, name):
    """Rebuild the scenario's model with the ORACLE from the same seeds; the stored checksums prove that the
    oracle's constructor consumes the RNG exactly like the reference's (same init order, same shapes)."""
    kw = dict(case["kwargs"])
    cls = O.VisionTransformerDistilled if name == "distilled" else O.VisionTransformer
    torch.manual_seed(case["seeds"]["init"])
    m = cls(**kw)
    RI.perturb(m, case["seeds"]["perturb"])
    sd = m.state_dict()
    assert list(sd.keys()) == case["state_keys"]
    check_sums(list(sd.items()), case["state_checksums"])
    return m


def run(model, x, loss_fn, seed):
    acts = []
    hooks = [b.register_forward_hook(lambda m, i, o: acts.append(o.detach().clone())) for b in model.blocks]
    if seed is not None:
        torch.manual_seed(seed)
    out = model(x)
    loss = loss_fn(out)
    loss.backward()
    for h in hooks:
        h.remove()
    return out, loss, acts


@pytest.mark.parametrize("name", ["avg", "token", "avg_ls_dp"])
def test_oracle_reproduces_reference_vit_micro(fx, name):
    """Block wiring (residual after DropPath(LayerScale(branch))), pooling, fc_norm/norm placement, LayerScale,
    the dpr rule and DropPath draw order: /root/reference/models/vision_transformer.py:80-178, 419-441, 452-623, 934-995."""
    case = fx["micro"][name]
    m = oracle_micro(case, name)
    if name == "avg":
        assert all(torch.equal(m.state_dict()[k], v) for k, v in case["state_dict"].items())
    m.train()
    masks = []
    for mod in m.modules():
        if isinstance(mod, O.DropPath):
            mod.register_forward_hook(lambda mm, i, o: masks.append((o[:, 0, 0] / i[0][:, 0, 0]).detach()))
    out, loss, acts = run(m, case["x"], lambda o: O.SoftTargetCrossEntropy()(o, case["target"]), case["seeds"]["fwd"])
    assert close(out, case["logits"]) and close(loss, case["loss"])
    assert len(acts) == len(case["acts"]) and all(close(a, b) for a, b in zip(acts, case["acts"]))
    assert len(masks) == len(case["masks"])
    for got, want in zip(masks, case["masks"]):
        assert torch.allclose(torch.nan_to_num(got, nan=0.0), want, atol=1e-5)
    for n, p in m.named_parameters():
        assert close(p.grad, case["grads"][n], 2e-6), n
    m.eval()
    with torch.no_grad():
        assert close(m(case["x"]), case["logits_eval"])


def test_oracle_reproduces_reference_distilled_micro(fx):
    """/root/reference/models/deit.py:28-119: two prefix tokens, two heads, (cls, dist) only when distilled_training and
    training, the average otherwise."""
    case = fx["micro"]["distilled"]
    m = oracle_micro(case, "distilled")
    m.train()
    m.set_distilled_training(True)
    labels, teacher = case["labels"], case["teacher"]
    out, loss, acts = run(m, case["x"], lambda o: 0.5 * F.cross_entropy(o[0], labels)
                          + 0.5 * F.cross_entropy(o[1], teacher.argmax(1)), None)
    assert isinstance(out, tuple) and close(out[0], case["logits"][0]) and close(out[1], case["logits"][1])
    assert close(loss, case["loss"]) and all(close(a, b) for a, b in zip(acts, case["acts"]))
    for n, p in m.named_parameters():
        assert close(p.grad, case["grads"][n], 2e-6), n
    m.set_distilled_training(False)
    assert close(m(case["x"]), case["logits_train_avg"])
    m.eval()
    with torch.no_grad():
        assert close(m(case["x"]), case["logits_eval"])


NAMED = ["vit_tiny_patch16_224", "vit_small_patch16_224", "vit_base_patch16_224", "vit_large_patch16_384",
         "deit_base_distilled_patch16_224", "deit_tiny_distilled_patch16_224", "my_vit_mini", "my_vit_ti", "my_vit_xs",
         "my_vit_s", "my_vit_b", "my_vit_l"]


@pytest.mark.parametrize("name", NAMED)
def test_entrypoints_layout_and_seeded_init_match_reference(fx, name):
    """The reference's own entrypoints (vision_transformer.py:2690-2860, my_vit.py:84-165, deit.py:306-314) constructed
    with the kwargs main.py:643-649 passes: state_dict keys/shapes, parameter count, per-block drop-path
    probabilities, img_size (384 for ViT-L/384 comes from the reference default_cfg through _builder.py:355-393), and
    the values of a seeded init (same RNG consumption order)."""
    rec = fx["named"][name]
    big = rec["n_params"] > 50_000_000
    if big:   # layout only: no need to draw 300 M random numbers
        with torch.device("meta"):
            m = O.create_model(name, **rec["kwargs"])
    else:
        torch.manual_seed(42)
        m = O.create_model(name, **rec["kwargs"])
    sd = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == rec["keys"]
    assert sum(p.numel() for p in m.parameters()) == rec["n_params"]
    assert tuple(m.patch_embed.img_size) == rec["img_size"]
    dp = [float(b.drop_path1.drop_prob) if hasattr(b.drop_path1, "drop_prob") else 0.0 for b in m.blocks]
    assert np.allclose(dp, rec["drop_probs"], atol=1e-7)
    if not big:
        check_sums(list(sd.items()), rec["init_checksums"])


def test_large_models_seeded_init_matches_reference(fx):
    """ViT-B: the headline config's init, value for value (checksums of all 152 tensors)."""
    rec = fx["named"]["vit_base_patch16_224"]
    torch.manual_seed(42)
    m = O.create_model("vit_base_patch16_224", **rec["kwargs"])
    check_sums(list(m.state_dict().items()), rec["init_checksums"])


@pytest.mark.parametrize("name", ["vit_tiny_patch16_224", "deit_tiny_distilled_patch16_224", "my_vit_mini", "my_vit_xs"])
def test_named_forward_backward_matches_reference(fx, name):
    """Config 1 (ViT-Ti, the CPU-runnable case), distilled DeiT-Ti and the head_dim 48 / 72 my_vit sizes with
    drop_path 0.1, on a seeded 2-image batch: logits, loss, DropPath masks, per-block activations, every gradient."""
    rec = fx["named"][name]
    x, tgt = RI.named_inputs()
    torch.manual_seed(42)
    m = O.create_model(name, **rec["kwargs"])
    m.train()
    ce = O.SoftTargetCrossEntropy()
    if name.startswith("deit_"):
        m.set_distilled_training(True)
        fn = lambda o: ce(o[0], tgt) + ce(o[1], tgt.flip(0))  # noqa: E731
    else:
        fn = lambda o: ce(o, tgt)  # noqa: E731
    out, loss, acts = run(m, x, fn, 11)
    outs = out if isinstance(out, tuple) else (out,)
    wants = rec["logits"] if isinstance(rec["logits"], tuple) else (rec["logits"],)
    assert all(close(a, b, 1e-5) for a, b in zip(outs, wants))
    assert close(loss, rec["loss"], 1e-6)
    assert close(acts[-1][:, :3, :16], rec["act_last_slice"], 1e-5)
    got = RI.checksums([(None, a) for a in acts])
    assert torch.allclose(got[:, 1], torch.stack(rec["act_checksums"])[:, 1], rtol=1e-5)
    grads = dict((n, p.grad) for n, p in m.named_parameters())
    assert list(grads.keys()) == rec["grad_keys"]
    gs = RI.checksums(list(grads.items()))
    assert torch.allclose(gs[:, 1], rec["grad_checksums"][:, 1], rtol=1e-4), "sum of squares of the gradients"
    for k, v in rec["grads_small"].items():
        assert close(grads[k], v, 1e-4), k


# ------------------------------------------------------------------------------------------------
# host logic: schedules, parameter groups, optimizer, KD loss / wrapper
# ------------------------------------------------------------------------------------------------
def test_cosine_scheduler_bit_for_bit(fx):
    """/root/reference/utils/__init__.py:667-684 -> the oracle's restatement AND the product's vectorised version."""
    from vision_transformers_torch_xla_b200 import utils as U

    for rec in fx["host"]["cosine_scheduler"]:
        for fn in (O.cosine_scheduler, U.cosine_scheduler):
            v = torch.from_numpy(np.asarray(fn(**rec["args"]), dtype=np.float64))
            assert v.numel() == rec["n"]
            assert torch.equal(v[::97], rec["every_97th"]), fn.__module__
            if rec["values"] is not None:
                assert torch.equal(v, rec["values"]), fn.__module__
            else:
                assert torch.allclose(RI.checksum(v), rec["checksum"], rtol=1e-13, atol=0)


def test_parameter_groups_match_reference(fx):
    """/root/reference/optim_factory.py:70-211: both the shape rule (non-TPU, :164) and the name rule (PJRT_DEVICE=TPU,
    :104-106) on a model with LayerScale — the one place where they differ (gamma)."""
    from vision_transformers_torch_xla_b200 import optim_factory as P

    torch.manual_seed(0)
    m = O.VisionTransformer(**dict(RI.MICRO, global_pool="avg", init_values=0.1))
    ids = {id(p): n for n, p in m.named_parameters()}
    assert sorted(m.no_weight_decay()) == fx["host"]["no_weight_decay"]

    def names(groups):
        return [dict(weight_decay=g["weight_decay"], lr_scale=g["lr_scale"], names=[ids[id(p)] for p in g["params"]])
                for g in groups]

    want = fx["host"]["param_groups"]
    assert names(O.get_parameter_groups(m, 0.05, m.no_weight_decay())) == want["shape"]
    assert names(O.get_parameter_groups(m, 0.05, m.no_weight_decay(), tpu_name_rule=True)) == want["tpu_name"]
    assert names(P.get_parameter_groups(m, 0.05, m.no_weight_decay())) == want["shape"]
    shape_nd = set(want["shape"][[g["weight_decay"] for g in want["shape"]].index(0.0)]["names"])
    name_nd = set(want["tpu_name"][[g["weight_decay"] for g in want["tpu_name"]].index(0.0)]["names"])
    assert {n for n in shape_nd - name_nd} == {f"blocks.{i}.ls{j}.gamma" for i in range(2) for j in (1, 2)}


def test_create_optimizer_matches_reference(fx):
    from vision_transformers_torch_xla_b200 import optim_factory as P

    want = fx["host"]["create_optimizer"]
    torch.manual_seed(0)
    m = O.VisionTransformer(**dict(RI.MICRO, global_pool="avg", init_values=0.1))
    opt = O.create_optimizer(m, lr=2e-3, weight_decay=0.05)
    assert type(opt).__name__ == want["cls"] == "AdamW"
    assert {k: opt.defaults[k] for k in ("lr", "betas", "eps", "weight_decay")} == want["defaults"]
    assert [g["weight_decay"] for g in opt.param_groups] == want["group_wd"]
    assert [len(g["params"]) for g in opt.param_groups] == want["group_sizes"]

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas = "adamw", 2e-3, 0.05, 1e-8, None

    mine = P.create_optimizer(Args, m)
    assert {k: mine.defaults[k] for k in ("lr", "betas", "eps", "weight_decay")} == want["defaults"]
    assert [g["weight_decay"] for g in mine.param_groups] == want["group_wd"]
    assert [len(g["params"]) for g in mine.param_groups] == want["group_sizes"]


def test_distillation_loss_matches_reference(fx):
    """/root/reference/main.py:939-968 (closure-local class): values and gradients for three (alpha, T) pairs, with a
    hard-label and a soft-label base criterion, and the tensor-input path."""
    kd = fx["host"]["kd"]
    s, t, y, ysoft = RI.kd_inputs()
    assert torch.equal(s, kd["student"]) and torch.equal(t, kd["teacher"])
    for (a, T), want in kd["cases"].items():
        s1 = s.clone().requires_grad_(True)
        l1 = O.DistillationLoss(nn.CrossEntropyLoss(), a, T)((s1, t), y)
        l1.backward()
        assert close(l1, want["hard_labels"][0]) and close(s1.grad, want["hard_labels"][1])
        s2 = s.clone().requires_grad_(True)
        l2 = O.DistillationLoss(O.SoftTargetCrossEntropy(), a, T)((s2, t), ysoft)
        l2.backward()
        assert close(l2, want["soft_labels"][0]) and close(s2.grad, want["soft_labels"][1])
        assert close(O.DistillationLoss(nn.CrossEntropyLoss(), a, T)(s, y), want["tensor_input"])


def test_distillation_wrapper_matches_reference(fx):
    """/root/reference/main.py:836-850: (student, teacher.no_grad) in train mode, student only in eval mode."""
    from vision_transformers_torch_xla_b200.losses import StudentWithDistillation as Mine

    want = fx["host"]["kd_wrapper"]
    for cls in (O.StudentWithDistillation, Mine):
        w = cls(nn.Linear(5, 3), nn.Linear(5, 3))
        x = torch.randn(2, 5)
        w.train()
        tr = w(x)
        w.eval()
        ev = w(x)
        got = dict(train_is_tuple=isinstance(tr, tuple), train_len=len(tr), teacher_requires_grad=bool(tr[1].requires_grad),
                   eval_is_tensor=torch.is_tensor(ev), state_keys=sorted(w.state_dict().keys()))
        assert got == want, cls


# ------------------------------------------------------------------------------------------------
# engine: the reference's own train_one_epoch / evaluate
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["uf1", "uf2"])
def test_oracle_engine_reproduces_reference_train_one_epoch(fx, name):
    """/root/reference/engine.py:19-333 executed as written (eager branch 257-274, schedule write 98-103 with the
    ``is not None`` weight-decay quirk, update_freq): per-micro-batch losses, returned meters, final weights and the
    final per-group lr / weight_decay over two epochs."""
    rec = fx["engine"][name]
    batches, _ = RI.engine_inputs()
    batches = batches[:rec["n_micro"]]
    check_sums([(None, b[0]) for b in batches], rec["batch_checksums"])
    m = O.VisionTransformer(**rec["kwargs"])
    m.load_state_dict(fx["micro"]["avg"]["state_dict"])
    opt = O.create_optimizer(m, **rec["args"])
    lr_s, wd_s = rec["lr_schedule"].numpy(), rec["wd_schedule"].numpy()
    spe, uf = rec["steps_per_epoch"], rec["update_freq"]
    half = rec["n_micro"] // 2
    losses = []
    for epoch in range(2):
        st = O.train_one_epoch(m, O.SoftTargetCrossEntropy(), batches[epoch * half:(epoch + 1) * half], opt, epoch,
                               start_steps=epoch * spe, lr_schedule_values=lr_s, wd_schedule_values=wd_s,
                               num_training_steps_per_epoch=spe, update_freq=uf)
        losses += st["losses"]
        assert abs(st["loss"] - rec["stats"][epoch]["loss"]) <= 1e-6 * abs(rec["stats"][epoch]["loss"])
        assert abs(st["lr"] - rec["stats"][epoch]["lr"]) <= 1e-12
    # the criterion saw the undivided loss; the engine logs loss / update_freq
    assert torch.allclose(torch.tensor(losses, dtype=torch.float64) * uf, rec["losses"].double(), rtol=2e-6)
    assert [g["weight_decay"] for g in opt.param_groups] == rec["final_group_wd"]   # the quirk: both groups decay
    assert [g["lr"] for g in opt.param_groups] == rec["final_group_lr"]
    sd = m.state_dict()
    assert list(sd.keys()) == rec["final_keys"]
    D = rec["kwargs"]["embed_dim"]
    for k, v in rec["final_small"].items():
        got_k = sd[k]
        if k.endswith("attn.qkv.bias"):
            # the key bias has an exactly-zero true gradient (softmax is invariant to a shift of all scores of a row);
            # what reaches Adam is rounding noise, which Adam normalises to +-lr steps: chaotic by construction
            got_k, v = torch.cat([got_k[:D], got_k[2 * D:]]), torch.cat([v[:D], v[2 * D:]])
        assert close(got_k, v, 5e-6), k
    got = RI.checksums(list(sd.items()))
    rows = [i for i, k in enumerate(sd.keys()) if not k.endswith("attn.qkv.bias")]
    assert torch.allclose(got[rows, 1], rec["final_checksums"][rows, 1], rtol=1e-5)


def test_oracle_evaluate_reproduces_reference(fx):
    """/root/reference/engine.py:339-430 on the model the uf1 run ends with."""
    rec = fx["engine"]["uf1"]
    batches, hard = RI.engine_inputs()
    m = O.VisionTransformer(**rec["kwargs"])
    m.load_state_dict(fx["micro"]["avg"]["state_dict"])
    opt = O.create_optimizer(m, **rec["args"])
    lr_s, wd_s = rec["lr_schedule"].numpy(), rec["wd_schedule"].numpy()
    for epoch in range(2):
        O.train_one_epoch(m, O.SoftTargetCrossEntropy(), batches[epoch * 5:(epoch + 1) * 5], opt, epoch,
                          start_steps=epoch * rec["steps_per_epoch"], lr_schedule_values=lr_s, wd_schedule_values=wd_s,
                          num_training_steps_per_epoch=rec["steps_per_epoch"], update_freq=1)
    got = O.evaluate(hard, m)
    want = fx["engine"]["evaluate"]["stats"]
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-5 * max(abs(want[k]), 1.0), (k, got[k], want[k])
