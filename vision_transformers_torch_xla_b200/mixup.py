"""Mixup / CutMix with label smoothing on the device — the `mixup_fn` the reference builds with
``timm.data.Mixup(mixup_alpha, cutmix_alpha, cutmix_minmax, prob, switch_prob, mode, label_smoothing, num_classes)``
(/root/reference/main.py:622-629) and applies to every batch inside the step (engine.py:259-262).

Same constructor, same call contract ``x, target = mixup_fn(x, target)``, same NumPy random draws in the same order as
timm 1.0.15 (so a seeded run mixes the same pairs with the same lam / box); the arithmetic is two vitk kernels
(`vitk_mixup_batch`, `vitk_mixup_target`) instead of ~8 eager ops with temporaries of the batch's size.
Built: mode='batch' (the reference default, `--mixup_mode batch`).  'elem' / 'pair' and `cutmix_minmax` raise."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def rand_bbox(img_shape, lam, margin=0.0, count=None):
    """timm.data.mixup.rand_bbox: square-ish box of area ratio (1 - lam), centre uniform over the image."""
    ratio = np.sqrt(1 - lam)
    img_h, img_w = img_shape[-2:]
    cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
    margin_y, margin_x = int(margin * cut_h), int(margin * cut_w)
    cy = np.random.randint(0 + margin_y, img_h - margin_y, size=count)
    cx = np.random.randint(0 + margin_x, img_w - margin_x, size=count)
    yl = np.clip(cy - cut_h // 2, 0, img_h)
    yh = np.clip(cy + cut_h // 2, 0, img_h)
    xl = np.clip(cx - cut_w // 2, 0, img_w)
    xh = np.clip(cx + cut_w // 2, 0, img_w)
    return yl, yh, xl, xh


def cutmix_bbox_and_lam(img_shape, lam, correct_lam=True, count=None):
    yl, yu, xl, xu = rand_bbox(img_shape, lam, count=count)
    if correct_lam:
        bbox_area = (yu - yl) * (xu - xl)
        lam = 1.0 - bbox_area / float(img_shape[-2] * img_shape[-1])
    return (yl, yu, xl, xu), lam


class Mixup:
    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch",
                 correct_lam=True, label_smoothing=0.1, num_classes=1000):
        if cutmix_minmax is not None:
            raise NotImplementedError("cutmix_minmax is not built (no reference launch script sets it)")
        if mode != "batch":
            raise NotImplementedError(f"mixup mode {mode!r}: only 'batch' (the reference default) is built")
        self.mixup_alpha = mixup_alpha
        self.cutmix_alpha = cutmix_alpha
        self.mix_prob = prob
        self.switch_prob = switch_prob
        self.label_smoothing = label_smoothing
        self.num_classes = num_classes
        self.mode = mode
        self.correct_lam = correct_lam
        self.mixup_enabled = True

    def _params_per_batch(self):
        lam = 1.0
        use_cutmix = False
        if self.mixup_enabled and np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = np.random.rand() < self.switch_prob
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else \
                    np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.0:
                lam_mix = np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.0:
                use_cutmix = True
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha)
            else:
                assert False, "One of mixup_alpha > 0., cutmix_alpha > 0., cutmix_minmax not None should be true."
            lam = float(lam_mix)
        return lam, use_cutmix

    def __call__(self, x: torch.Tensor, target: torch.Tensor):
        assert len(x) % 2 == 0, "Batch size should be even when using this"
        if not x.is_cuda:
            raise L.VitkError("Mixup runs on the device: move the batch to the GPU first (engine.train_one_epoch does)")
        lam, use_cutmix = self._params_per_batch()
        if lam != 1.0:
            x = x.contiguous()
            if use_cutmix:
                box, lam = cutmix_bbox_and_lam(x.shape, lam, correct_lam=self.correct_lam)
                L.mixup_batch(x, lam, True, box)
            else:
                L.mixup_batch(x, lam, False)
        out = torch.empty(x.shape[0], self.num_classes, dtype=torch.float32, device=x.device)
        L.mixup_target(target.to(x.device, torch.int64).contiguous(), out, lam, self.label_smoothing)
        return x, out
