"""Training losses of the reference (train.py:18-53, utils.py:286-299), kept in PyTorch as in the
reference: they act on the [m,3] / [n,3] logits only and define what flows into the backward kernels."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


@torch.no_grad()
def labels_to_balanced_weights(labels, merge_lu=True):
    """Inverse-frequency class weights; for two-sided problems the lower/upper weights are averaged
    (utils.py:286-299)."""
    # torch.unique(return_counts) of the reference == per-class counts; computed with a comparison + sum so that
    # there is neither a device sort nor a host sync (unique and bincount both read a size back to the host)
    cls = torch.arange(3, device=labels.device)
    cnt = (labels.view(-1, 1) == cls).sum(0).to(torch.float32)
    present = cnt > 0
    res = torch.where(present, cnt.sum() / cnt.clamp_min(1.0), torch.zeros_like(cnt))
    if merge_lu:
        merged = (res[0] + res[2]) / 2.
        two_sided = present.sum() != 2            # the reference merges lower/upper unless exactly 2 classes occur
        res = torch.where(two_sided & (cls != 1), merged, res)     # (no host-built mask: a pageable H2D copy would sync)
    return res


class FocalLoss(nn.modules.loss._WeightedLoss):
    """train.py:18-28 (note: the focal factor is applied to the batch-mean CE, as in the reference)."""

    def __init__(self, weight=None, gamma=2, reduction="mean"):
        super().__init__(weight, reduction=reduction)
        self.gamma = gamma
        self.weight = weight

    def forward(self, input, target):
        ce = F.cross_entropy(input, target, reduction=self.reduction, weight=self.weight)
        pt = torch.exp(-ce)
        return ((1 - pt) ** self.gamma * ce).mean()


cri_focal = FocalLoss()


def unbalanced(logpr_cons, logpr_vars, y_s, y_t):
    return F.cross_entropy(torch.cat((logpr_cons, logpr_vars), dim=0), torch.cat((y_s, y_t), dim=0))


def balanced(logpr_cons, logpr_vars, y_s, y_t):
    m, n = len(y_s), len(y_t)
    loss = (m + n) / m * F.cross_entropy(logpr_cons, y_s, weight=labels_to_balanced_weights(y_s))
    loss = loss + (m + n) / n * F.cross_entropy(logpr_vars, y_t, weight=labels_to_balanced_weights(y_t))
    return loss


def focal(logpr_cons, logpr_vars, y_s, y_t):
    return cri_focal(torch.cat((logpr_cons, logpr_vars), dim=0), torch.cat((y_s, y_t), dim=0))


LOSSES = {"balanced": balanced, "unbalanced": unbalanced, "focal": focal}
