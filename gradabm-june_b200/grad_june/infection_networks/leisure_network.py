"""Leisure networks (reference: grad_june/infection_networks/leisure_network.py).

All leisure classes run on the single ``leisure`` edge set; what differs is the per-agent attendance
probability table [day type, sex, age] that scales both transmissions and susceptibilities, and for
care visits the extra (age > 75) factor on the susceptible side.  The kernels read the table through
the agent's (sex, age) class byte, so no per-agent probability vectors are materialised.
"""
import torch

from .. import ops
from ..utils import parse_age_probabilities
from .base import InfectionNetwork


class LeisureNetwork(InfectionNetwork):
    kind = ops.KIND_LEISURE

    def __init__(self, log_beta, leisure_probabilities, device):
        super().__init__(log_beta=log_beta, device=device)
        self.leisure_probabilities = self._parse_leisure_probabilities(leisure_probabilities)
        self.weekday_probabilities = None
        self.weekend_probabilities = None

    @classmethod
    def from_parameters(cls, params):
        name = cls._get_name()
        return cls(device=params["system"]["device"], leisure_probabilities=params["leisure"][name],
                   **params["networks"][name])

    def _parse_leisure_probabilities(self, leisure_probabilities):
        table = torch.zeros((2, 2, 100), device=self.device)
        for i, day_type in enumerate(("weekday", "weekend")):
            for j, sex in enumerate(("male", "female")):
                table[i, j, :] = torch.tensor(parse_age_probabilities(leisure_probabilities[day_type][sex]),
                                              device=self.device)
        return table

    def initialize_leisure_probabilities(self, data):
        """Per-agent vectors, kept for API parity; the kernels index the table directly."""
        sex, age = data["agent"].sex, data["agent"].age
        self.weekday_probabilities = self.leisure_probabilities[0, sex, age]
        self.weekend_probabilities = self.leisure_probabilities[1, sex, age]

    def edge_type(self):
        return "leisure"

    def _agent_mask(self, data, policies, timer):
        """quarantine mask x attendance probability by (day type, sex, age) — leisure_network.py:61-85."""
        if self.weekday_probabilities is None:
            self.initialize_leisure_probabilities(data)
        prob = self.weekday_probabilities if timer.day_type == "weekday" else self.weekend_probabilities
        return super()._agent_mask(data, policies, timer) * prob


class PubNetwork(LeisureNetwork):
    pass


class CinemaNetwork(LeisureNetwork):
    pass


class GroceryNetwork(LeisureNetwork):
    pass


class GymNetwork(LeisureNetwork):
    pass


class VisitNetwork(LeisureNetwork):
    pass


class CareVisitNetwork(LeisureNetwork):
    kind = ops.KIND_CARE_VISIT   # susceptibilities additionally masked by age > 75

    def _get_susceptibilities(self, data, policies, timer):
        return super()._get_susceptibilities(data, policies, timer) * (data["agent"].age > 75)
