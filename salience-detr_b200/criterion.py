"""Training-side companions of the path: the salience supervision of the filter's score maps and a
loader for the reference's checkpoints.

``SalienceCriterion`` mirrors the reference class (models/detectors/salience_detr.py:13-116): same constructor arguments,
same ``forward(foreground_mask, targets, feature_strides, image_sizes) -> {"loss_salience": ...}``.  The (image, level)
target maps -- for every token the best scale-independent centredness over the ground-truth boxes that contain its pixel
centre and belong to its level -- come from ONE kernel (``sdetr_salience_targets``) instead of (HW, boxes, 4) tensors per image
and level; the focal loss over the (b, Nv) logits stays in torch so that it is differentiable w.r.t. the score maps
(``sigmoid_focal_loss``, models/bricks/losses.py:4-12).

``load_reference_checkpoint`` accepts what the reference writes / publishes (util/utils.py:358-422, README "Model zoo"):
a bare ``state_dict`` or ``{"model": state_dict, ...}``, with or without DDP's ``module.`` prefix, for the whole detector
(``transformer.*`` keys are picked) or for the transformer alone; shape-mismatched entries are skipped like the reference's
``filter_mismatched_weights``."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
from torch import Tensor, nn
from torch.nn import functional as F

from . import cabi


def sigmoid_focal_loss(inputs: Tensor, targets: Tensor, num_boxes, alpha: float = 0.25, gamma: float = 2) -> Tensor:
    """models/bricks/losses.py:4-12 (soft targets; the weight keeps its gradient)."""
    prob = inputs.sigmoid()
    weight = (1 - alpha) * prob ** gamma * (1 - targets) + targets * alpha * (1 - prob) ** gamma
    loss = F.binary_cross_entropy_with_logits(inputs, targets.to(inputs.dtype), reduction="none") * weight
    return (loss.sum(1) / max(loss.shape[1], 1)).sum() / num_boxes


class SalienceCriterion(nn.Module):
    def __init__(self, limit_range: Tuple = ((-1, 64), (64, 128), (128, 256), (256, 99999)), noise_scale: float = 0.0,
                 alpha: float = 0.25, gamma: float = 2.0):
        super().__init__()
        self.limit_range = limit_range
        self.noise_scale = noise_scale
        self.alpha = alpha
        self.gamma = gamma

    @torch.no_grad()
    def mask_targets(self, shapes: Sequence[Tuple[int, int]], targets: List[Dict], feature_strides, image_sizes, device) -> Tensor:
        """(b, Nv) target maps (salience_detr.py:28-46, 64-114)."""
        boxes_px = []
        for t, (img_h, img_w) in zip(targets, image_sizes):
            cx, cy, w, h = t["boxes"].to(device=device, dtype=torch.float32).unbind(-1)
            xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)  # box_ops._box_cxcywh_to_xyxy
            boxes_px.append(xyxy * torch.tensor([img_w, img_h, img_w, img_h], device=device))
        if device.type == "cuda":
            mb = max(1, max(bx.shape[0] for bx in boxes_px))
            padded = torch.zeros(len(boxes_px), mb, 4, device=device)
            for i, bx in enumerate(boxes_px):
                padded[i, :bx.shape[0]] = bx
            counts = torch.tensor([bx.shape[0] for bx in boxes_px], dtype=torch.int32, device=device)
            tgt = cabi.salience_targets(padded, counts, list(shapes), [tuple(float(v) for v in s) for s in feature_strides],
                                        self.limit_range)
        else:  # host-side restatement (CPU tests): same arithmetic in torch
            per_level = []
            for lvl, ((h, w), (sy, sx)) in enumerate(zip(shapes, feature_strides)):
                ys = (torch.arange(h, dtype=torch.float32) + 0.5) * sy
                xs = (torch.arange(w, dtype=torch.float32) + 0.5) * sx
                cy, cx = [g.reshape(-1) for g in torch.meshgrid(ys, xs, indexing="ij")]
                rows = []
                for bx in boxes_px:
                    if bx.shape[0] == 0:
                        rows.append(torch.zeros(h * w))
                        continue
                    dl, dt = cx[:, None] - bx[None, :, 0], cy[:, None] - bx[None, :, 1]
                    dr, db = bx[None, :, 2] - cx[:, None], bx[None, :, 3] - cy[:, None]
                    d = torch.stack([dl, dt, dr, db], -1)
                    inside = d.amin(-1) > 0
                    lo, hi = self.limit_range[lvl]
                    pos = (inside & (d.amax(-1) > lo) & (d.amax(-1) <= hi)).any(-1)
                    conf = 1 - torch.sqrt(((dl - dr) / (dl + dr)) ** 2 + ((dt - db) / (dt + db)) ** 2) / 2
                    best = torch.where(inside, conf, torch.zeros_like(conf)).amax(-1)
                    rows.append(torch.where(pos, best, torch.zeros_like(best)))
                per_level.append(torch.stack(rows))
            tgt = torch.cat(per_level, 1)
        if self.noise_scale:
            tgt = (1 - self.noise_scale) * tgt + self.noise_scale * torch.rand_like(tgt)
        return tgt

    def forward(self, foreground_mask, targets, feature_strides, image_sizes):
        shapes = [tuple(m.shape[-2:]) for m in foreground_mask]
        logits = torch.cat([e.flatten(-2) for e in foreground_mask], -1).squeeze(1)
        mask_targets = self.mask_targets(shapes, targets, feature_strides, image_sizes, logits.device)
        num_pos = torch.sum(mask_targets > 0.5 * self.noise_scale).clamp_(min=1)
        loss = sigmoid_focal_loss(logits, mask_targets, num_pos, alpha=self.alpha, gamma=self.gamma) * logits.shape[1]
        return {"loss_salience": loss}


def load_reference_checkpoint(model: nn.Module, checkpoint, prefix: str = "transformer.") -> Dict[str, list]:
    """Load a reference checkpoint (path or mapping) into one of this package's modules.  -> dict(missing=..., unexpected=...,
    mismatched=...).  Keys are matched after stripping DDP's ``module.`` and -- when the checkpoint is a whole detector --
    the ``prefix`` of the sub-module."""
    if isinstance(checkpoint, (str, bytes)):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    sd = checkpoint.get("model", checkpoint) if isinstance(checkpoint, dict) else checkpoint
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items() if torch.is_tensor(v)}
    if prefix and any(k.startswith(prefix) for k in sd):
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    own = model.state_dict()
    mismatched = [k for k, v in sd.items() if k in own and own[k].shape != v.shape]
    for k in mismatched:  # util/utils.py:358-367: keep the model's value for mismatched shapes
        sd[k] = own[k]
    res = model.load_state_dict(sd, strict=False)
    return {"missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys), "mismatched": mismatched}
