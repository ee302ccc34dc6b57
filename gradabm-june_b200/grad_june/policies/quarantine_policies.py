"""Quarantine (grad_june/policies/quarantine_policies.py:13-33).

``apply`` keeps the reference contract (it materialises ``quarantine_mask`` as a float tensor, which
the stand-alone network modules read); the fused step kernels instead take the scalar thresholds of
the active policies (``active_thresholds``) and evaluate ``stage < threshold`` in registers.
"""
import torch

from .policies import Policy, PolicyCollection


class Quarantine(Policy):
    spec = "quarantine"

    def __init__(self, start_date, end_date, stage_threshold, device):
        super().__init__(start_date=start_date, end_date=end_date, device=device)
        self.stage_threshold = stage_threshold

    def apply(self, symptom_stages, timer):
        if self.is_active(timer.date):
            return (symptom_stages < self.stage_threshold).to(torch.float)
        return torch.ones(symptom_stages.shape, device=symptom_stages.device)


class QuarantinePolicies(PolicyCollection):
    def __init__(self, policies):
        super().__init__(policies)
        self.quarantine_mask = 1.0

    def active_thresholds(self, timer):
        return [float(p.stage_threshold) for p in self.policies if p.is_active(timer.date)]

    def apply(self, symptom_stages, timer):
        mask = torch.ones(symptom_stages.shape, device=symptom_stages.device)
        for policy in self.policies:
            mask = mask * policy.apply(symptom_stages=symptom_stages, timer=timer)
        self.quarantine_mask = mask
