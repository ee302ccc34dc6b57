"""Infection draw and state-update helpers (reference: grad_june/infection.py).

``IsInfectedSampler`` is the Gumbel-softmax (tau = 0.1, hard, straight-through) Bernoulli draw of
infection.py:3-18, executed by the SAMPLE phase of ``gj_step_forward`` with counter-based Philox
noise (or injected noise inside ``ops.inject_noise``).
"""
import torch

from . import ops
from .world import DeviceWorld, build_csr

_PLAIN_WORLDS = {}


def _agent_only_world(n, device):
    """Descriptor with agents but no edges, for stand-alone phases that need no graph."""
    key = (n, str(device))
    w = _PLAIN_WORLDS.get(key)
    if w is None:
        if len(_PLAIN_WORLDS) > 8:
            _PLAIN_WORLDS.clear()
        zeros = torch.zeros(n, dtype=torch.long, device=device)
        w = build_csr(n, [], {}, {}, {}, zeros, zeros, 16, 1024, device)
        _PLAIN_WORLDS[key] = w
    return w


class IsInfectedSampler(torch.nn.Module):
    def forward(self, not_infected_probs):
        ops.require_cuda(not_infected_probs, "not_infected_probs")
        world = _agent_only_world(not_infected_probs.numel(), not_infected_probs.device)
        spec = ops.StepSpec(now=0.0, dt=0.0, day_type=0, nets=[], quarantine=None, phases=ops.PHASE_SAMPLE,
                            want_reductions=False)
        out = ops.infection_step(ops.StepStatic(world=world), spec, None, {}, q_in=not_infected_probs)
        return out["n"]


def infect_people(data, timer, new_infected):
    """Seeding variant of the state update: clamp (gradient 1 at the boundary) — infection.py:21-28."""
    agent = data["agent"]
    agent.susceptibility = torch.clamp(agent.susceptibility - new_infected, min=0.0)
    agent.is_infected = agent.is_infected + new_infected
    agent.infection_time = agent.infection_time + new_infected * (timer.now - agent.infection_time)


def infect_fraction_of_people(data, timer, symptoms_updater, fraction, device):
    n_agents = data["agent"].id.shape[0]
    probs = fraction * torch.ones(n_agents, device=device)
    new_infected = IsInfectedSampler()(1.0 - probs)
    infect_people(data, timer, new_infected)
    return new_infected


def infect_people_at_indices(data, indices, device="cpu"):
    """Deterministic seeding used by tests (infection.py:45-63): infected at t=0, stage exposed next."""
    agent = data["agent"]
    sym = agent["symptoms"]
    idx = torch.as_tensor(indices, dtype=torch.long)

    def patched(t, value):
        t = t.detach().cpu().clone()
        t[idx] = value
        return t.to(device)

    agent["susceptibility"] = patched(agent["susceptibility"], 0.0)
    agent["is_infected"] = patched(agent["is_infected"], 1.0)
    agent["infection_time"] = patched(agent["infection_time"], 0.0)
    sym["next_stage"] = patched(sym["next_stage"], 2)
    sym["current_stage"] = patched(sym["current_stage"], 1)
    return data
