"""Drop-in for the contrastive losses of the reference's `src/clip/train/losses.py` (forward values).

`InfoNCELoss(temperature)(features_a, features_b)` and `JointContrastiveLoss(...)(image, query, target)` keep the
reference's names, argument order, defaults and return shape `(loss, metrics)`.  The reference builds the (B, B) logits
`A @ B.T / T` and calls `F.cross_entropy` on them and on their transpose (`losses.py:45-55`); here one fused kernel
per direction (`kemr_infonce_rows`) keeps a row of A in registers, walks B out of L2 and folds max / sum-of-exp on
the fly, so the logits never reach HBM.  fp32 arithmetic like torch; the two agree to a few ulp of the loss.

Gradients: the returned loss carries a backward that re-forms the softmax blocks with torch matmuls (training the
encoders is outside the hot path of this engine; the forward value is what validation loops log).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib, engine


def _row_losses(a: torch.Tensor, b: torch.Tensor, temperature: float) -> torch.Tensor:
    engine._require_cuda()
    a = a.detach().to(device="cuda", dtype=torch.float32).contiguous()
    b = b.detach().to(device="cuda", dtype=torch.float32).contiguous()
    if a.dim() != 2 or a.shape != b.shape:
        raise _lib.KemrError(f"InfoNCE needs two (B, D) feature matrices, got {tuple(a.shape)} and {tuple(b.shape)}")
    out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().kemr_infonce_rows(engine._ptr(a), engine._ptr(b), a.shape[0], a.shape[1], float(temperature),
                                             engine._ptr(out), engine._stream()))
    return out


class _InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fa, fb, temperature):
        la = _row_losses(fa, fb, temperature).mean()
        lb = _row_losses(fb, fa, temperature).mean()
        ctx.save_for_backward(fa, fb)
        ctx.temperature = temperature
        return (la + lb) / 2.0, la, lb

    @staticmethod
    def backward(ctx, g, _ga, _gb):
        fa, fb = ctx.saved_tensors
        B = fa.shape[0]
        logits = (fa.float() @ fb.float().T) / ctx.temperature
        eye = torch.eye(B, device=logits.device, dtype=logits.dtype)
        gl = (torch.softmax(logits, dim=1) - eye + (torch.softmax(logits.T, dim=1) - eye).T) / (2.0 * B) * g
        return (gl @ fb.float() / ctx.temperature).to(fa.dtype), (gl.T @ fa.float() / ctx.temperature).to(fb.dtype), None


class InfoNCELoss:
    """Reference `losses.py:11-64`: symmetric InfoNCE, (loss_a2b + loss_b2a) / 2."""

    def __init__(self, temperature: float = 0.07):
        self.temperature = temperature

    def forward(self, features_a: torch.Tensor, features_b: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, float]]:
        fa, fb = features_a.cuda(), features_b.cuda()
        loss, la, lb = _InfoNCEFn.apply(fa, fb, float(self.temperature))
        return loss, {"loss": loss.item(), "loss_a2b": la.item(), "loss_b2a": lb.item()}

    __call__ = forward


class JointContrastiveLoss:
    """Reference `losses.py:67-140`: t2i_weight * InfoNCE(target, image) + t2t_weight * InfoNCE(query, target), the
    weights normalised to sum to one."""

    def __init__(self, temperature: float = 0.07, t2i_weight: float = 0.5, t2t_weight: float = 0.5):
        self.temperature = temperature
        self.infonce = InfoNCELoss(temperature=temperature)
        weight_sum = t2i_weight + t2t_weight
        self.t2i_weight = t2i_weight / weight_sum
        self.t2t_weight = t2t_weight / weight_sum

    def forward(self, image_features, query_features, target_features) -> Tuple[torch.Tensor, Dict[str, float]]:
        loss_t2i, _ = self.infonce(target_features, image_features)
        loss_t2t, _ = self.infonce(query_features, target_features)
        total = self.t2i_weight * loss_t2i + self.t2t_weight * loss_t2t
        return total, {"loss": total.item(), "loss_t2i": loss_t2i.item(), "loss_t2t": loss_t2t.item(),
                       "t2i_weight": self.t2i_weight, "t2t_weight": self.t2t_weight}

    __call__ = forward
