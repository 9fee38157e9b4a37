"""Optimizer factory (mirrors /root/reference/optim_factory.py:70-296 for the path the configs use:
``--opt adamw`` with the decay / no_decay parameter-group rule) on the fused flat AdamW kernel."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .store import ALIGN, ParamStore


def get_parameter_groups(model: nn.Module, weight_decay: float = 1e-5, skip_list=(), get_num_layer=None,
                         get_layer_scale=None) -> List[dict]:
    """Two groups, ``decay`` / ``no_decay``, each carrying ``lr_scale`` — the non-TPU shape rule of
    /root/reference/optim_factory.py:155-195: no decay iff ``ndim == 1 or name.endswith('.bias') or name in skip``."""
    groups: Dict[str, dict] = {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith(".bias") or name in skip_list:
            group_name, this_wd = "no_decay", 0.0
        else:
            group_name, this_wd = "decay", weight_decay
        layer_id = None
        if get_num_layer is not None:
            layer_id = get_num_layer(name)
            group_name = "layer_%d_%s" % (layer_id, group_name)
        if group_name not in groups:
            scale = get_layer_scale(layer_id) if get_layer_scale is not None else 1.0
            groups[group_name] = {"weight_decay": this_wd, "params": [], "lr_scale": scale}
        groups[group_name]["params"].append(param)
    return list(groups.values())


class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` semantics, one kernel launch per step per model replica.

    Parameters must live in a ``ParamStore`` (they do after the model's first forward, or after
    ``store.get_store(model)``).  ``param_groups[i]['lr' | 'weight_decay' | 'lr_scale']`` are read at every
    ``step()`` exactly like the reference's engine writes them (engine.py:98-103).  The same launch refreshes
    the bf16 weight shadow, optionally zeroes the gradients (``zero_grad`` then costs nothing), folds in the
    1/world_size of data-parallel gradient averaging (``grad_scale``) and an EMA of the weights."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, fused_zero_grad=True):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid AdamW hyper-parameters")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.fused_zero_grad = fused_zero_grad
        self.grad_scale = 1.0
        self.ema: Optional[torch.Tensor] = None
        self.ema_decay = 0.0
        self._plan = None
        self._step = 0
        self._grads_zeroed = False
        #: callbacks(optimizer) that complete the gradient (data-parallel all-reduce); run ONCE per step, by
        #: ``sync_gradients()`` — called from ``step()``, or earlier by gradient clipping, which needs the reduced gradient
        self.pre_step_hooks = []
        self._synced = False
        #: set by the data-parallel layer for one step: bf16 copy of the summed gradient that the kernel reads instead
        self.grad_lowp: Optional[torch.Tensor] = None
        #: set by ``engine.clip_grad_norm_`` for one step: device scalar multiplied into the gradient (clip coefficient)
        self.grad_scale_dev: Optional[torch.Tensor] = None

    def sync_gradients(self) -> None:
        """Run the pre-step hooks (gradient all-reduce) if they have not run for this step yet."""
        if not self._synced:
            for hook in self.pre_step_hooks:
                hook(self)
            self._synced = True

    # ---- plan: which store, which chunk belongs to which group ----
    def _build_plan(self):
        old_plan, old_ema = self._plan, self.ema
        # moments that already exist (a plan rebuilt after model.to() / reset_classifier, or a state loaded before the
        # store existed) are carried over by parameter instead of being silently reset
        carried = {id(p): (st_["exp_avg"].detach().clone(), st_["exp_avg_sq"].detach().clone())
                   for p, st_ in self.state.items() if "exp_avg" in st_ and "exp_avg_sq" in st_}
        stores: Dict[int, ParamStore] = {}
        owner: Dict[int, int] = {}
        for gi, group in enumerate(self.param_groups):
            if gi >= 8:
                raise NotImplementedError("FusedAdamW supports at most 8 parameter groups")
            b1, b2 = group["betas"]
            if (b1, b2, group["eps"]) != (self.param_groups[0]["betas"][0], self.param_groups[0]["betas"][1],
                                          self.param_groups[0]["eps"]):
                raise NotImplementedError("FusedAdamW: betas/eps must be shared by all groups")
            for p in group["params"]:
                owner[id(p)] = gi
        plan = []
        seen = set()
        for group in self.param_groups:
            for p in group["params"]:
                st = getattr(p, "_vitk_store_ref", None)
                st = st() if st is not None else None
                if st is not None and not st.valid():
                    st = None
                if st is None or id(p) not in st.offsets:
                    raise L.VitkError("FusedAdamW: a parameter is not in a vitk ParamStore. Run one forward pass "
                                      "(or store.get_store(model)) before optimizer.step(); there is no eager fallback")
                if id(st) not in seen:
                    seen.add(id(st))
                    stores[id(st)] = st
        for st in stores.values():
            nchunk = st.total // ALIGN
            table = torch.zeros(nchunk, dtype=torch.uint8)
            frozen = torch.ones(nchunk, dtype=torch.bool)
            for p in st.params:
                o, n = st.offsets[id(p)]
                c0, c1 = o // ALIGN, (o + n + ALIGN - 1) // ALIGN
                if id(p) in owner:
                    table[c0:c1] = owner[id(p)]
                    frozen[c0:c1] = False
            if bool(frozen.any()):
                # parameters outside this optimizer (e.g. frozen): give them a group with lr = wd = 0
                ng = len(self.param_groups)
                if ng >= 8:
                    raise NotImplementedError("no spare group slot for frozen parameters")
                table[frozen] = ng
            m = torch.zeros(st.total, dtype=torch.float32, device=st.device)
            v = torch.zeros(st.total, dtype=torch.float32, device=st.device)
            plan.append(dict(store=st, table=table.to(st.device), m=m, v=v, has_frozen=bool(frozen.any())))
            for p in st.params:
                if id(p) in owner:
                    o, n = st.offsets[id(p)]
                    self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": m[o:o + n].view(p.shape),
                                     "exp_avg_sq": v[o:o + n].view(p.shape)}
                    if id(p) in carried and carried[id(p)][0].numel() == n:
                        self.state[p]["exp_avg"].copy_(carried[id(p)][0].to(m.device).view(p.shape))
                        self.state[p]["exp_avg_sq"].copy_(carried[id(p)][1].to(v.device).view(p.shape))
        self._plan = plan
        if old_ema is not None and old_plan is not None:
            # the fused EMA follows the parameters into the new flat layout (new parameters start from their weights)
            if len(plan) != 1:
                raise NotImplementedError("EMA with several stores")
            new_st, old_st = plan[0]["store"], old_plan[0]["store"]
            ema = new_st.flat.clone()
            for p in new_st.params:
                if id(p) in old_st.offsets and old_st.offsets[id(p)][1] == new_st.offsets[id(p)][1]:
                    (oo, n), (no, _) = old_st.offsets[id(p)], new_st.offsets[id(p)]
                    ema[no:no + n].copy_(old_ema[oo:oo + n].to(ema.device))
            self.ema = ema

    def enable_ema(self, decay: float):
        """Keep ``ema = decay*ema + (1-decay)*p`` inside the AdamW launch (timm ModelEma semantics)."""
        if self._plan is None:
            self._build_plan()
        if len(self._plan) != 1:
            raise NotImplementedError("EMA with several stores")
        self.ema = self._plan[0]["store"].flat.clone()
        self.ema_decay = decay
        return self.ema

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._plan is None or any(not e["store"].valid() for e in self._plan):
            self._build_plan()
        self.sync_gradients()
        self._synced = False
        self._step += 1
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        if self.grad_lowp is not None and len(self._plan) != 1:
            raise NotImplementedError("a low-precision gradient copy with several stores")
        for e in self._plan:
            st = e["store"]
            lrs = [float(g["lr"]) for g in self.param_groups]
            wds = [float(g["weight_decay"]) for g in self.param_groups]
            if e["has_frozen"]:
                lrs.append(0.0)
                wds.append(0.0)
            L.adamw_flat(st.flat, st.grad, e["m"], e["v"], st.shadow, self.ema, e["table"], ALIGN, lrs, wds, b1, b2, eps,
                         self._step, grad_scale=self.grad_scale, ema_decay=self.ema_decay,
                         zero_grad=self.fused_zero_grad, g_bf16=self.grad_lowp, grad_scale_dev=self.grad_scale_dev)
            st.mark_shadow_current()
        self.grad_lowp = None
        self.grad_scale_dev = None
        self._grads_zeroed = self.fused_zero_grad
        for p_state in self.state.values():
            p_state["step"] = torch.tensor(float(self._step))
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """Gradients are views of the flat buffer: keep them, zero in place (free right after ``step``)."""
        if self._plan is None:
            return super().zero_grad(set_to_none=False)
        if self._grads_zeroed:
            self._grads_zeroed = False
            return
        for e in self._plan:
            e["store"].grad.zero_()

    def load_state_dict(self, state_dict):
        """Works before the model's first forward too (the reference resume order: create model -> create optimizer ->
        auto_load_model): if the parameters are not in a ParamStore yet, the loaded moments stay in ``self.state`` and
        are moved into the flat buffers when the plan is built."""
        if self._plan is None:
            try:
                self._build_plan()
            except L.VitkError:
                self._plan = None
        views = {id(p): dict(s) for p, s in self.state.items()} if self._plan is not None else {}
        super().load_state_dict(state_dict)
        steps = []
        for p, s in list(self.state.items()):
            old = views.get(id(p))
            if old is None:
                continue
            for k in ("exp_avg", "exp_avg_sq"):
                if k in s and s[k].data_ptr() != old[k].data_ptr():
                    old[k].copy_(s[k])
                    s[k] = old[k]
            if "step" in s:
                steps.append(int(float(s["step"])))
        if steps:
            self._step = max(steps)


def create_optimizer(args, model, get_num_layer=None, get_layer_scale=None, filter_bias_and_bn=True, skip_list=None):
    """Same call contract as /root/reference/optim_factory.py:214-296.  Only ``adamw`` (what every reference
    config and launch script uses, main.py:175-186) is built; other names raise NotImplementedError."""
    opt_lower = args.opt.lower()
    weight_decay = args.weight_decay
    if filter_bias_and_bn:
        skip = {}
        if skip_list is not None:
            skip = skip_list
        elif hasattr(model, "no_weight_decay"):
            skip = model.no_weight_decay()
        parameters = get_parameter_groups(model, weight_decay, skip, get_num_layer, get_layer_scale)
        weight_decay = 0.0
    else:
        parameters = [p for p in model.parameters() if p.requires_grad]
    opt_args = dict(lr=args.lr, weight_decay=weight_decay)
    if getattr(args, "opt_eps", None) is not None:
        opt_args["eps"] = args.opt_eps
    if getattr(args, "opt_betas", None) is not None:
        opt_args["betas"] = tuple(args.opt_betas)
    opt_split = opt_lower.split("_")
    opt_lower = opt_split[-1]
    if opt_lower in ("adamw", "fusedadamw"):
        optimizer = FusedAdamW(parameters, **opt_args)
    else:
        raise NotImplementedError(f"optimizer {args.opt!r}: only adamw is built on the B200 path "
                                  "(no reference config uses another one)")
    if len(opt_split) > 1 and opt_split[0] == "lookahead":
        raise NotImplementedError("lookahead is not built")
    return optimizer
