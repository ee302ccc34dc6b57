"""Venue closure: drops networks from the step's activity list (close_venue_policies.py:11-22)."""
from .policies import Policy, PolicyCollection


class CloseVenue(Policy):
    spec = "close_venue"

    def __init__(self, start_date, end_date, names, device):
        super().__init__(start_date=start_date, end_date=end_date, device=device)
        self.edge_type_to_close = {f"{name}" for name in names}

    def apply(self, edge_types, timer):
        if not self.is_active(timer.date):
            return edge_types
        return [e for e in edge_types if e not in self.edge_type_to_close]


class CloseVenuePolicies(PolicyCollection):
    def apply(self, edge_types, timer):
        for policy in self.policies:
            edge_types = policy.apply(edge_types=edge_types, timer=timer)
        return edge_types
