"""Losses of the hot path on the fused CE kernel (vitk_ce_fwd_bwd).

``SoftTargetCrossEntropy`` / ``LabelSmoothingCrossEntropy`` are the timm losses the reference selects at
/root/reference/main.py:926-935 (import path there: ``timm.loss``); ``DistillationLoss`` and
``StudentWithDistillation`` are the closure-local classes of /root/reference/main.py:836-850, 939-968
(duplicated in test_kd.py:43-88), published here at module level with the same constructor and forward
contracts.  ``hard=True`` adds the DeiT hard-label variant needed by the distilled student (config 4)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .ops import CEFn


def _check_logits(x):
    if x.dim() != 2:
        raise ValueError(f"expected [batch, classes] logits, got shape {tuple(x.shape)}")
    return x.float() if x.dtype != torch.float32 else x


class SoftTargetCrossEntropy(nn.Module):
    """``mean_b sum_c -t[b,c] log_softmax(x)[b,c]`` — forward and dlogits in one kernel."""

    def forward(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return CEFn.apply(_check_logits(x), target, None, 0.0, None, 0.0, 1.0)


class LabelSmoothingCrossEntropy(nn.Module):
    """NLL loss with label smoothing: ``(1-eps) nll + eps mean_c(-logp)``."""

    def __init__(self, smoothing: float = 0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing = smoothing
        self.confidence = 1.0 - smoothing

    def forward(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return CEFn.apply(_check_logits(x), None, target, self.smoothing, None, 0.0, 1.0)


class CrossEntropyLoss(nn.Module):
    """``torch.nn.CrossEntropyLoss()`` (mean reduction) for hard int64 labels or soft float targets."""

    def __init__(self, label_smoothing: float = 0.0):
        super().__init__()
        self.label_smoothing = label_smoothing

    def forward(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.dtype in (torch.int64, torch.int32):
            return CEFn.apply(_check_logits(x), None, target, self.label_smoothing, None, 0.0, 1.0)
        if self.label_smoothing:
            raise NotImplementedError("label_smoothing with soft targets is not built")
        return CEFn.apply(_check_logits(x), target, None, 0.0, None, 0.0, 1.0)


def _base_kind(base_criterion):
    """(soft?, smoothing) of a base criterion the fused kernel can absorb, else None."""
    if isinstance(base_criterion, SoftTargetCrossEntropy):
        return True, 0.0
    if isinstance(base_criterion, LabelSmoothingCrossEntropy):
        return False, base_criterion.smoothing
    if isinstance(base_criterion, CrossEntropyLoss):
        return None, base_criterion.label_smoothing
    if isinstance(base_criterion, nn.CrossEntropyLoss) and base_criterion.reduction == "mean" \
            and base_criterion.weight is None and base_criterion.ignore_index == -100:
        return None, base_criterion.label_smoothing
    return None


class DistillationLoss(nn.Module):
    """``(1-alpha) * base(student, y) + alpha * T^2 * KLDiv_batchmean(log_softmax(s/T), softmax(t/T))``.

    ``outputs`` is either a tensor (-> base criterion only) or ``(student_logits, teacher_logits)``; with a
    distilled DeiT student ``student_logits`` may itself be ``(cls_logits, dist_logits)``: the base loss is
    taken on the class head, the distillation term on the distillation head.  ``hard=True`` replaces the KL
    term by CE against ``argmax(teacher)`` (DeiT hard distillation)."""

    def __init__(self, base_criterion, alpha: float = 0.7, temperature: float = 4.0, hard: bool = False):
        super().__init__()
        self.base_criterion = base_criterion
        self.alpha = alpha
        self.temperature = temperature
        self.hard = hard

    def forward(self, outputs, targets):
        if not isinstance(outputs, tuple):
            return self.base_criterion(outputs, targets)
        student_logits, teacher_logits = outputs
        dist_logits = None
        if isinstance(student_logits, tuple):
            student_logits, dist_logits = student_logits
        teacher_logits = teacher_logits.detach()
        kind = _base_kind(self.base_criterion)
        if kind is None:
            raise NotImplementedError(f"DistillationLoss: base criterion {type(self.base_criterion).__name__} is not "
                                      "one of the fused losses")
        soft_flag, smoothing = kind
        is_soft = targets.dtype.is_floating_point if soft_flag is None else soft_flag
        soft = targets if is_soft else None
        labels = None if is_soft else targets
        s = _check_logits(student_logits)
        if self.hard:
            ce = CEFn.apply(s, soft, labels, smoothing, None, 0.0, 1.0)
            d = _check_logits(dist_logits if dist_logits is not None else student_logits)
            kd = CEFn.apply(d, None, teacher_logits.argmax(dim=1), 0.0, None, 0.0, 1.0)
            return (1 - self.alpha) * ce + self.alpha * kd
        if dist_logits is None:
            # single launch: base + KD on the same logits
            return CEFn.apply(s, soft, labels, smoothing, teacher_logits.float(), self.alpha, self.temperature)
        ce = CEFn.apply(s, soft, labels, smoothing, None, 0.0, 1.0)
        # pure KD term on the distillation head: alpha=1 makes the kernel return T^2 * KL only
        kd = CEFn.apply(_check_logits(dist_logits), None, torch.zeros(s.shape[0], dtype=torch.int64, device=s.device),
                        0.0, teacher_logits.float(), 1.0, self.temperature)
        return (1 - self.alpha) * ce + self.alpha * kd


class StudentWithDistillation(nn.Module):
    """Returns ``(student_logits, teacher_logits)`` in training mode (teacher under no_grad), else the
    student logits only — /root/reference/main.py:836-850."""

    def __init__(self, student_model, teacher_model):
        super().__init__()
        self.student = student_model
        self.teacher = teacher_model

    def forward(self, x):
        student_logits = self.student(x)
        if self.training and self.teacher is not None:
            with torch.no_grad():
                teacher_logits = self.teacher(x)
            return student_logits, teacher_logits
        return student_logits
