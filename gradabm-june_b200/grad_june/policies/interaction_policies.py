"""Beta multipliers (grad_june/policies/interaction_policies.py:10-31)."""
import torch

from .policies import Policy, PolicyCollection


class InteractionPolicy(Policy):
    spec = "interaction"


class InteractionPolicies(PolicyCollection):
    def apply(self, beta, name, timer):
        for policy in self.policies:
            beta = policy.apply(beta=beta, name=name, timer=timer)
        return beta


class SocialDistancing(InteractionPolicy):
    """While active, scales a network's beta by ``beta_factors[name]``, else ``["all"]``, else 1."""

    def __init__(self, start_date, end_date, beta_factors, device):
        super().__init__(start_date=start_date, end_date=end_date, device=device)
        self.beta_factors = {k: torch.tensor(float(v), device=device) for k, v in beta_factors.items()}

    def factor(self, name):
        if name in self.beta_factors:
            return self.beta_factors[name]
        return self.beta_factors.get("all", torch.tensor(1.0))

    def apply(self, beta, name, timer):
        if not self.is_active(timer.date):
            return beta
        return beta * self.factor(name)
