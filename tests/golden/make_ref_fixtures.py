"""Generates tests/golden/ref_fixtures.pt by EXECUTING THE REFERENCE'S OWN CODE (this container only).

The reference (/root/reference) cannot be imported as a package offline: every module pulls in ``timm`` /
``torch_xla`` / ``tensorflow`` at import time.  But the logic of the hot path that lives IN the
reference tree is timm-free once the imports are out of the way, so this script

  1. parses the reference sources with ``ast`` and pulls out the definitions of the hot path, verbatim, at run
     time (nothing is copied into this repository):
        models/vision_transformer.py : LayerScale, Block, global_pool_nlc, VisionTransformer, init_weights_vit_*,
                                       get_init_weights_vit, _cfg, default_cfgs, _create_vision_transformer,
                                       the vit_{tiny,small,base,large}_patch16_{224,384} entrypoints
        models/deit.py               : VisionTransformerDistilled, _create_deit, deit_*_distilled_patch16_224
        models/my_vit.py             : my_vit_{mini,ti,xs,s,b,l} and their cfg helpers
        models/_manipulate.py        : named_apply
        models/_builder.py           : _update_default_model_kwargs, _filter_kwargs
        utils/__init__.py            : SmoothedValue, MetricLogger, cosine_scheduler, get_rank, ...
        optim_factory.py             : get_parameter_groups, create_optimizer
        engine.py                    : train_one_epoch, evaluate
        main.py (inside main())      : StudentWithDistillation, DistillationLoss
  2. ``exec``s them in namespaces where the names the reference resolves from pip ``timm`` through
     models/_compat.py:27-172 (Attention, Mlp, PatchEmbed, LayerNorm, DropPath, trunc_normal_, accuracy, ...) are
     bound to the ORACLE's restatements of those leaf layers (timm itself is absent) — so everything that is the
     reference's own code (block wiring, residual order, LayerScale, pooling, constructor wiring, init order,
     entrypoint sizes, distilled head, param-group rule, schedules, KD loss/wrapper, the engine's eager branch
     with its weight-decay quirk) runs as written;
  3. runs seeded scenarios and stores their outputs as small fixtures.

tests/test_ref_fixtures.py then checks the oracle (and the product's host logic) against these outputs on CPU and
tests/test_gpu_ref_fixtures.py checks the CUDA path against them on the B200.

Run from the repo root (needs /root/reference):   python tests/golden/make_ref_fixtures.py
"""
from __future__ import annotations

import ast
import copy
import math
import os
import sys
import types
from collections import OrderedDict, defaultdict, deque
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import vit_oracle as O  # noqa: E402  (leaf layers only: the timm names of models/_compat.py)

REF = os.environ.get("VITK_REFERENCE", "/root/reference")


# ------------------------------------------------------------------------------------------------
# ast extraction
# ------------------------------------------------------------------------------------------------
def _source(path):
    with open(os.path.join(REF, path), encoding="utf-8") as f:
        return f.read()


def extract(path, names, ns, inside=None, assigns=()):
    """exec the top-level (or, with ``inside='fn'``, nested-in-that-function) definitions ``names`` and the
    top-level assignments to ``assigns`` of reference file ``path`` in namespace ``ns``, in source order."""
    src = _source(path)
    tree = ast.parse(src)
    want, want_assign = set(names), set(assigns)
    found = set()

    def run(node):
        exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, path), "exec"), ns)

    if inside is not None:
        outer = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == inside)
        nodes = [n for n in ast.walk(outer) if n is not outer]
    else:
        nodes = tree.body
    for node in nodes:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in want:
            run(node)
            found.add(node.name)
        elif isinstance(node, ast.Assign) and inside is None:
            tgt = [t.id for t in node.targets if isinstance(t, ast.Name)]
            if tgt and tgt[0] in want_assign:
                run(node)
                found.add(tgt[0])
    missing = (want | want_assign) - found
    if missing:
        raise RuntimeError(f"{path}: definitions not found: {sorted(missing)}")
    return ns


def quiet_print(*a, **k):
    pass


# ------------------------------------------------------------------------------------------------
# the timm names of models/_compat.py, bound to the oracle's restated leaf layers
# ------------------------------------------------------------------------------------------------
class Attention(O.Attention):
    """Constructor signature the reference Block uses (vision_transformer.py:149-159)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, scale_norm=False, proj_bias=True,
                 attn_drop=0.0, proj_drop=0.0, norm_layer=None):
        assert not qk_norm and not scale_norm
        super().__init__(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                         proj_drop=proj_drop)


class Mlp(O.Mlp):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                 bias=True, drop=0.0):
        assert act_layer is nn.GELU and norm_layer is None
        super().__init__(in_features, hidden_features, out_features, bias=bias, drop=drop)


class PatchEmbed(O.PatchEmbed):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, bias=True,
                 dynamic_img_pad=False, **kw):
        assert norm_layer is None and not dynamic_img_pad and not kw
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, bias=bias)


MASKS = []   # DropPath masks in draw order (so that a CUDA run can replay them)


class DropPath(O.DropPath):
    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep_prob = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
        if keep_prob > 0.0 and self.scale_by_keep:
            random_tensor.div_(keep_prob)
        MASKS.append(random_tensor.flatten().clone())
        return x * random_tensor


def accuracy(output, target, topk=(1,)):
    """timm.utils.accuracy 1.0.15."""
    maxk = min(max(topk), output.size()[1])
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[:min(k, maxk)].reshape(-1).float().sum(0) * 100. / batch_size for k in topk]


def build_namespaces():
    import logging
    import typing

    base = dict(torch=torch, nn=nn, F=F, math=math, os=os, copy=copy, partial=partial, OrderedDict=OrderedDict,
                logging=logging, print=quiet_print, np=np, defaultdict=defaultdict, deque=deque)
    for n in ("Any", "Callable", "Dict", "Optional", "Set", "Tuple", "Type", "Union", "List", "Iterable", "Literal"):
        base[n] = getattr(typing, n)
    base["Final"] = torch.jit.Final

    # ---- models/vision_transformer.py ----
    vt = dict(base)
    vt.update(Attention=Attention, Mlp=Mlp, PatchEmbed=PatchEmbed, LayerNorm=O.LayerNorm, DropPath=DropPath,
              trunc_normal_=O.trunc_normal_, get_norm_layer=lambda x: x, get_act_layer=lambda x: x,
              lecun_normal_=None, AttentionPoolLatent=None, PatchDropout=None, checkpoint_filter_fn=lambda *a, **k: None,
              register_model=lambda f: f, generate_default_cfgs=lambda d: d,
              IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
              IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5), IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5),
              OPENAI_CLIP_MEAN=(0.48145466, 0.4578275, 0.40821073), OPENAI_CLIP_STD=(0.26862954, 0.26130258, 0.27577711))
    extract("models/_manipulate.py", ["named_apply"], vt)
    extract("models/_builder.py", ["_update_default_model_kwargs", "_filter_kwargs"], vt)

    def build_model_with_cfg(model_cls, variant, pretrained, pretrained_filter_fn=None, pretrained_strict=True,
                             feature_cfg=None, **kwargs):
        """Stand-in for models/_builder.py:432-551 without the pip-timm registry: looks the variant up in the
        extracted ``default_cfgs`` and lets the reference's own ``_update_default_model_kwargs`` (355-393) fill
        num_classes / global_pool / in_chans / img_size, then constructs ``model_cls(**kwargs)`` (493)."""
        assert not pretrained
        for k in ("pretrained_cfg", "pretrained_cfg_overlay", "cache_dir"):
            kwargs.pop(k, None)
        cfgs = build_model_with_cfg.default_cfgs
        cfg = next((c for k, c in cfgs.items() if k.split(".")[0] == variant), None)
        if cfg is not None:
            vt["_update_default_model_kwargs"](dict(cfg), kwargs, None)
        return model_cls(**kwargs)

    vt["build_model_with_cfg"] = build_model_with_cfg
    entry = [f"vit_{s}_patch16_{r}" for s in ("tiny", "small", "base", "large") for r in (224, 384)]
    extract("models/vision_transformer.py",
            ["LayerScale", "Block", "global_pool_nlc", "VisionTransformer", "init_weights_vit_timm",
             "init_weights_vit_jax", "init_weights_vit_moco", "get_init_weights_vit", "_cfg",
             "_create_vision_transformer"] + entry, vt, assigns=["default_cfgs", "_USE_NAFLEX_DEFAULT"])
    build_model_with_cfg.default_cfgs = vt["default_cfgs"]

    # ---- models/deit.py (base class: the in-tree VisionTransformer extracted above) ----
    de = dict(base)
    de.update(VisionTransformer=vt["VisionTransformer"], trunc_normal_=O.trunc_normal_,
              checkpoint_filter_fn=lambda *a, **k: None,
              resample_abs_pos_embed=None, register_model=lambda f: f, generate_default_cfgs=lambda d: d,
              IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))

    def build_deit(model_cls, variant, pretrained, pretrained_filter_fn=None, feature_cfg=None, **kwargs):
        assert not pretrained
        cfg = next((c for k, c in de["default_cfgs"].items() if k.split(".")[0] == variant), None)
        if cfg is not None:
            vt["_update_default_model_kwargs"](dict(cfg), kwargs, None)
        return model_cls(**kwargs)

    de["build_model_with_cfg"] = build_deit
    extract("models/deit.py", ["VisionTransformerDistilled", "_create_deit", "_cfg", "deit_tiny_distilled_patch16_224",
                               "deit_base_distilled_patch16_224"], de, assigns=["default_cfgs"])

    # ---- models/my_vit.py ----
    my = dict(base)
    my.update(VisionTransformer=vt["VisionTransformer"], _create_vision_transformer=vt["_create_vision_transformer"],
              default_cfgs=vt["default_cfgs"], register_model=lambda f: f)
    import dataclasses
    my.update(asdict=dataclasses.asdict, is_dataclass=dataclasses.is_dataclass)
    extract("models/my_vit.py", ["_cfg_to_dict", "_resolve_cfg", "_apply_default_cfg", "my_vit_mini", "my_vit_ti",
                                 "my_vit_xs", "my_vit_s", "my_vit_b", "my_vit_l"], my, assigns=["_CFG_LOOKUP"])

    # ---- utils/__init__.py ----
    import datetime
    import time

    ut = dict(base)
    ut.update(time=time, datetime=datetime, dist=torch.distributed)
    extract("utils/__init__.py", ["SmoothedValue", "MetricLogger", "is_dist_avail_and_initialized", "get_world_size",
                                  "get_rank", "is_main_process", "cosine_scheduler"], ut)
    utils_mod = types.SimpleNamespace(**{k: ut[k] for k in ("SmoothedValue", "MetricLogger", "get_rank", "get_world_size",
                                                            "is_main_process", "cosine_scheduler")})

    # ---- optim_factory.py ----
    import json

    of = dict(base)
    of.update(optim=torch.optim, json=json, has_apex=False)
    extract("optim_factory.py", ["get_parameter_groups", "create_optimizer"], of)

    # ---- engine.py ----
    en = dict(base)
    en.update(utils=utils_mod, time=time, Mixup=object, ModelEma=object, accuracy=accuracy)
    extract("engine.py", ["train_one_epoch", "evaluate"], en)

    # ---- main.py: the closure-local KD classes ----
    mn = dict(base)
    extract("main.py", ["StudentWithDistillation", "DistillationLoss"], mn, inside="main")
    return types.SimpleNamespace(vt=vt, de=de, my=my, ut=ut, of=of, en=en, mn=mn, utils=utils_mod)


# ------------------------------------------------------------------------------------------------
# scenarios
# ------------------------------------------------------------------------------------------------
from ref_inputs import (MICRO, SMALL, checksum, checksums, engine_inputs, kd_inputs, micro_inputs,  # noqa: E402
                        named_inputs, perturb)


def run_model(model, x, loss_fn, seed_fwd=None):
    """forward (per-block activations via hooks) + loss + backward; returns everything as plain tensors."""
    acts = []
    hooks = [blk.register_forward_hook(lambda m, i, o: acts.append(o.detach().clone())) for blk in model.blocks]
    MASKS.clear()
    if seed_fwd is not None:
        torch.manual_seed(seed_fwd)
    model.zero_grad(set_to_none=True)
    out = model(x)
    loss = loss_fn(out)
    loss.backward()
    for h in hooks:
        h.remove()
    outs = tuple(o.detach().clone() for o in out) if isinstance(out, tuple) else out.detach().clone()
    return dict(logits=outs, loss=loss.detach().clone(), acts=acts, masks=[m.clone() for m in MASKS],
                grads={n: p.grad.detach().clone() for n, p in model.named_parameters()})


def soft_ce(t):
    return lambda out: torch.sum(-t * F.log_softmax(out, dim=-1), dim=-1).mean()   # timm SoftTargetCrossEntropy


def micro_cases(R):
    VT, VTD = R.vt["VisionTransformer"], R.de["VisionTransformerDistilled"]
    cases = {}
    x, tgt, labels, teacher = micro_inputs()
    for name, kw in (("avg", dict(global_pool="avg")), ("token", dict(global_pool="token")),
                     ("avg_ls_dp", dict(global_pool="avg", init_values=0.1, drop_path_rate=0.3, depth=4))):
        torch.manual_seed(1234)
        model = VT(**dict(MICRO, **kw))
        perturb(model, 7)
        model.train()
        r = run_model(model, x, soft_ce(tgt), seed_fwd=4321)
        sd = model.state_dict()
        r.update(kwargs=dict(MICRO, **kw), x=x, target=tgt, seeds=dict(init=1234, perturb=7, fwd=4321),
                 state_keys=list(sd.keys()), state_checksums=checksums(sd.items()))
        if name == "avg":   # 'token' shares every weight (fc_norm.* <-> norm.*); the others are rebuilt from the seeds
            r["state_dict"] = {k: v.clone() for k, v in sd.items()}
        model.eval()
        with torch.no_grad():
            r["logits_eval"] = model(x).clone()
        cases[name] = r
    # distilled (deit.py): train + distilled_training -> (cls, dist); DeiT hard-label loss on the tuple
    torch.manual_seed(1234)
    model = VTD(**dict(MICRO, global_pool="token"))
    perturb(model, 7)
    model.train()
    model.set_distilled_training(True)
    hard = lambda out: 0.5 * F.cross_entropy(out[0], labels) + 0.5 * F.cross_entropy(out[1], teacher.argmax(1))  # noqa: E731
    r = run_model(model, x, hard)
    sd = model.state_dict()
    r.update(kwargs=dict(MICRO, global_pool="token"), x=x, labels=labels, teacher=teacher,
             seeds=dict(init=1234, perturb=7, fwd=None), state_keys=list(sd.keys()), state_checksums=checksums(sd.items()))
    model.set_distilled_training(False)
    r["logits_train_avg"] = model(x).detach().clone()
    model.eval()
    with torch.no_grad():
        r["logits_eval"] = model(x).clone()
    cases["distilled"] = r
    return cases


def named_cases(R):
    """The five BASELINE configs + my_vit sizes through the reference's own entrypoints: state_dict layout,
    parameter counts, seeded-init checksums; ViT-Ti / DeiT-Ti also run forward + backward."""
    out = {}
    ctor = dict(vit_tiny_patch16_224=R.vt, vit_small_patch16_224=R.vt, vit_base_patch16_224=R.vt,
                vit_large_patch16_384=R.vt, deit_base_distilled_patch16_224=R.de, deit_tiny_distilled_patch16_224=R.de,
                my_vit_mini=R.my, my_vit_ti=R.my, my_vit_xs=R.my, my_vit_s=R.my, my_vit_b=R.my, my_vit_l=R.my)
    x, tgt = named_inputs()
    for name, ns in ctor.items():
        kw = dict(num_classes=1000, drop_path_rate=0.1)
        if not name.startswith("deit_"):
            kw["global_pool"] = "avg"   # what main.py:643-649 passes
        torch.manual_seed(42)
        model = ns[name](pretrained=False, **kw)
        sd = model.state_dict()
        rec = dict(kwargs=kw, keys=[(k, tuple(v.shape)) for k, v in sd.items()],
                   n_params=sum(p.numel() for p in model.parameters()),
                   init_checksums=checksums(sd.items()),
                   drop_probs=[(float(b.drop_path1.drop_prob) if hasattr(b.drop_path1, "drop_prob") else 0.0)
                               for b in model.blocks],
                   img_size=tuple(model.patch_embed.img_size))
        if name in ("vit_tiny_patch16_224", "deit_tiny_distilled_patch16_224", "my_vit_mini", "my_vit_xs"):
            model.train()
            if name.startswith("deit_"):
                model.set_distilled_training(True)
                fn = lambda o: soft_ce(tgt)(o[0]) + soft_ce(tgt.flip(0))(o[1])  # noqa: E731
            else:
                fn = soft_ce(tgt)
            r = run_model(model, x, fn, seed_fwd=11)
            rec.update(logits=r["logits"], loss=r["loss"], masks=r["masks"],
                       act_checksums=[checksum(a) for a in r["acts"]],
                       act_last_slice=r["acts"][-1][:, :3, :16].clone(),
                       grad_keys=list(r["grads"].keys()), grad_checksums=checksums(r["grads"].items()),
                       grads_small={k: v for k, v in r["grads"].items() if v.numel() <= 1024})
        out[name] = rec
        print(f"  {name}: {rec['n_params']} params, {len(rec['keys'])} tensors", file=sys.stderr)
        del model
    return out


def host_cases(R):
    out = {}
    cs = R.ut["cosine_scheduler"]
    sched_args = [dict(base_value=4e-3, final_value=1e-6, epochs=3, niter_per_ep=7, warmup_epochs=1),
                  dict(base_value=0.05, final_value=0.05, epochs=3, niter_per_ep=7),
                  dict(base_value=1e-3, final_value=1e-5, epochs=5, niter_per_ep=11, warmup_epochs=2, warmup_steps=9,
                       start_warmup_value=1e-6),
                  dict(base_value=4e-3, final_value=1e-6, epochs=300, niter_per_ep=312, warmup_epochs=20)]
    scheds = []
    for a in sched_args:
        v = torch.from_numpy(np.asarray(cs(**a), dtype=np.float64))
        scheds.append(dict(args=a, n=v.numel(), checksum=checksum(v), values=v if v.numel() <= 4096 else None,
                           every_97th=v[::97].clone()))
    out["cosine_scheduler"] = scheds

    # parameter groups (optim_factory.py:70-211) under both rules, on a model with LayerScale
    VT = R.vt["VisionTransformer"]
    torch.manual_seed(0)
    model = VT(**dict(MICRO, global_pool="avg", init_values=0.1))
    ids = {id(p): n for n, p in model.named_parameters()}
    groups = {}
    for rule, env in (("shape", None), ("tpu_name", "TPU")):
        old = os.environ.pop("PJRT_DEVICE", None)
        if env:
            os.environ["PJRT_DEVICE"] = env
        try:
            gs = R.of["get_parameter_groups"](model, 0.05, model.no_weight_decay())
        finally:
            os.environ.pop("PJRT_DEVICE", None)
            if old is not None:
                os.environ["PJRT_DEVICE"] = old
        groups[rule] = [dict(weight_decay=g_["weight_decay"], lr_scale=g_["lr_scale"],
                             names=[ids[id(p)] for p in g_["params"]]) for g_ in gs]
    out["param_groups"] = groups
    out["no_weight_decay"] = sorted(model.no_weight_decay())

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas, momentum = "adamw", 2e-3, 0.05, 1e-8, None, 0.9

    opt = R.of["create_optimizer"](Args, model)
    out["create_optimizer"] = dict(cls=type(opt).__name__, defaults={k: opt.defaults[k] for k in ("lr", "betas", "eps", "weight_decay")},
                                   group_wd=[g_["weight_decay"] for g_ in opt.param_groups],
                                   group_sizes=[len(g_["params"]) for g_ in opt.param_groups])

    # KD loss + wrapper (main.py:836-850, 939-968)
    DL, SW = R.mn["DistillationLoss"], R.mn["StudentWithDistillation"]
    s, t, y, ysoft = kd_inputs()
    g = torch.Generator().manual_seed(31)
    kd = {}
    for a, T in ((0.7, 4.0), (0.3, 1.0), (1.0, 2.5)):
        s1 = s.clone().requires_grad_(True)
        l1 = DL(nn.CrossEntropyLoss(), a, T)((s1, t), y)
        l1.backward()
        s2 = s.clone().requires_grad_(True)
        l2 = DL(O.SoftTargetCrossEntropy(), a, T)((s2, t), ysoft)
        l2.backward()
        kd[(a, T)] = dict(hard_labels=(l1.detach(), s1.grad.clone()), soft_labels=(l2.detach(), s2.grad.clone()),
                          tensor_input=DL(nn.CrossEntropyLoss(), a, T)(s, y).detach())
    out["kd"] = dict(student=s, teacher=t, labels=y, soft=ysoft, cases=kd)

    stu, tea = nn.Linear(5, 3), nn.Linear(5, 3)
    w = SW(stu, tea)
    xin = torch.randn(2, 5, generator=g)
    w.train()
    tr = w(xin)
    w.eval()
    ev = w(xin)
    out["kd_wrapper"] = dict(train_is_tuple=isinstance(tr, tuple), train_len=len(tr) if isinstance(tr, tuple) else 0,
                             teacher_requires_grad=bool(tr[1].requires_grad), eval_is_tensor=torch.is_tensor(ev),
                             state_keys=sorted(w.state_dict().keys()))
    return out


def engine_cases(R):
    """The reference's own train_one_epoch (eager fp32 branch, engine.py:257-274, schedule write 98-103) and evaluate
    (339-430) on the micro model: per-step losses, returned meters, final weights."""
    VT = R.vt["VisionTransformer"]
    out = {}
    batches, hard_batches = engine_inputs()

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas, momentum = "adamw", 2e-3, 0.05, 1e-8, None, 0.9

    for name, update_freq, n_micro in (("uf1", 1, 10), ("uf2", 2, 12)):
        torch.manual_seed(1234)
        model = VT(**dict(MICRO, global_pool="avg"))
        perturb(model, 7)
        opt = R.of["create_optimizer"](Args, model)
        steps = n_micro // update_freq
        lr_s = R.ut["cosine_scheduler"](2e-3, 1e-5, 2, steps // 2, warmup_epochs=1)
        wd_s = R.ut["cosine_scheduler"](0.05, 0.1, 2, steps // 2)
        losses = []

        class Crit(nn.Module):
            def forward(self, o, t):
                loss = torch.sum(-t * F.log_softmax(o, dim=-1), dim=-1).mean()
                losses.append(loss.detach().clone())
                return loss

        stats = {}
        for epoch in range(2):
            lo = epoch * (n_micro // 2)
            stats[epoch] = R.en["train_one_epoch"](
                model, Crit(), batches[lo:lo + n_micro // 2], opt, torch.device("cpu"), epoch, None, max_norm=0,
                model_ema=None, mixup_fn=lambda s_, t_: (s_, t_), start_steps=epoch * (steps // 2), lr_schedule_values=lr_s,
                wd_schedule_values=wd_s, num_training_steps_per_epoch=steps // 2, update_freq=update_freq, use_amp=False,
                tpu=False, log_freq=1)
        fsd = model.state_dict()
        out[name] = dict(kwargs=dict(MICRO, global_pool="avg"), init="micro/avg/state_dict", args=dict(lr=2e-3, weight_decay=0.05),
                         n_micro=n_micro, batch_checksums=checksums([(None, b[0]) for b in batches[:n_micro]]),
                         update_freq=update_freq, lr_schedule=torch.from_numpy(lr_s),
                         wd_schedule=torch.from_numpy(wd_s), steps_per_epoch=steps // 2, losses=torch.stack(losses),
                         stats=stats, final_keys=list(fsd.keys()), final_checksums=checksums(fsd.items()),
                         final_small={k: v.clone() for k, v in fsd.items() if v.numel() <= SMALL},
                         final_group_wd=[g_["weight_decay"] for g_ in opt.param_groups],
                         final_group_lr=[g_["lr"] for g_ in opt.param_groups])
        if name == "uf1":
            ev = R.en["evaluate"](hard_batches, model, torch.device("cpu"), use_amp=False, tpu=False)
            out["evaluate"] = dict(stats=ev, note="3 hard-label batches drawn right after the 12 soft ones from the same generator")
    return out


def main():
    torch.set_num_threads(1)          # deterministic summation order
    torch.use_deterministic_algorithms(True)
    R = build_namespaces()
    fx = dict(meta=dict(torch=torch.__version__, reference=REF,
                        note="outputs of the reference's own in-tree code executed with oracle leaf layers; see the "
                             "docstring of tests/golden/make_ref_fixtures.py"))
    fx["micro"] = micro_cases(R)
    fx["host"] = host_cases(R)
    fx["engine"] = engine_cases(R)
    fx["named"] = named_cases(R)
    path = os.path.join(HERE, "ref_fixtures.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
