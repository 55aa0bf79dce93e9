"""torch.nn.CTCLoss-compatible callable backed by the CTC kernels (train.py:196 call contract)."""
import os

import torch

from .functions import CTCLossFn


class CTCLoss(torch.nn.Module):
    """CTCLoss(blank, reduction='none'|'mean'|'sum', zero_infinity=False).

    forward(log_probs [T, N, C], targets [N, S], input_lengths [N], target_lengths [N]) -> [N] (reduction='none').
    The reference uses reduction='none' then torch.mean (train.py:77-78).
    """

    def __init__(self, blank=0, reduction="mean", zero_infinity=False):
        super().__init__()
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        dev = log_probs.device
        if targets.dim() != 2:
            raise NotImplementedError("lightning_asr_b200 CTCLoss expects padded 2-D targets [N, S] (train.py:246-247)")
        targets = targets.to(dev).long().contiguous()
        il = torch.as_tensor(input_lengths).to(dev).int().contiguous()
        tl = torch.as_tensor(target_lengths).to(dev).int().contiguous()
        if os.environ.get("LASR_DEBUG", "0") == "1":  # host-side validation costs a device sync: debug runs only
            bad = (targets < 0) | (targets >= log_probs.shape[-1])
            valid = torch.arange(targets.shape[1], device=dev)[None, :] < tl[:, None]
            if bool((bad & valid).any()):
                raise ValueError("CTCLoss: a target label is outside [0, C)")
            if bool((tl > targets.shape[1]).any()) or bool((il > log_probs.shape[0]).any()):
                raise ValueError("CTCLoss: a length exceeds its tensor")
        nll = CTCLossFn.apply(log_probs, targets, il, tl, self.blank, self.zero_infinity)
        if self.zero_infinity:
            nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
        if self.reduction == "none":
            return nll
        if self.reduction == "sum":
            return nll.sum()
        return (nll / tl.clamp_min(1).to(nll.dtype)).mean()  # torch's 'mean': per-target-length, then batch mean
