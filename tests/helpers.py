"""Shared test plumbing: build oracle inputs from the golden fixtures and from the package's own
host-side logic (Timer / Policies / network table), so that the CPU oracle and the CUDA path are
driven by the same schedule."""
import copy
import json
from pathlib import Path

import numpy as np
import torch

import grad_june
from grad_june import Timer
from grad_june.infection_networks import InfectionNetworks
from grad_june.policies import Policies
from grad_june.symptoms import SymptomsSampler
from noise import make_noise
from oracle import gj_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
SAMPLE_TYPES = ["household", "company", "school", "university", "care_home", "leisure"]


def load_params(tag):
    """Parameters of a golden run, as recorded by make_golden.py (dates arrive as strings, int keys as
    strings: normalise back to what yaml.safe_load gives)."""
    with open(GOLDEN / f"schedule_{tag}.json") as f:
        blob = json.load(f)
    params = blob["params"]

    def fix(d):
        if isinstance(d, dict):
            out = {}
            for k, v in d.items():
                if isinstance(k, str) and k.lstrip("-").isdigit():
                    k = int(k)
                out[k] = fix(v)
            return out
        if isinstance(d, list):
            return [fix(x) for x in d]
        return d
    return fix(params), blob["schedule"]


def oracle_world(arrays, types, device="cpu"):
    w = O.OracleWorld(n_agents=len(arrays["age"]),
                      age=torch.as_tensor(np.asarray(arrays["age"]).astype(np.int64), device=device),
                      sex=torch.as_tensor(np.asarray(arrays["sex"]).astype(np.int64), device=device))
    for t in types:
        w.edges[t] = O.EdgeType(
            src=torch.as_tensor(np.asarray(arrays[f"{t}_src"]).astype(np.int64), device=device),
            dst=torch.as_tensor(np.asarray(arrays[f"{t}_dst"]).astype(np.int64), device=device),
            people=torch.as_tensor(np.asarray(arrays[f"{t}_people"]).astype(np.int64), device=device),
            n_groups=int(arrays[f"{t}_ngroups"]))
    return w


def oracle_symptoms(sampler: SymptomsSampler, device="cpu"):
    return O.SymptomsSpec(n_stages=len(sampler.stages),
                          prob=sampler.stage_transition_probabilities.detach().to(device),
                          trans_times=dict(sampler._host_times[0]), rec_times=dict(sampler._host_times[1]))


def oracle_schedule(params, networks: InfectionNetworks, policies: Policies, device="cpu"):
    """StepSpec list from the package's host logic; betas stay attached to the networks' log_beta."""
    timer = Timer.from_parameters(params)
    steps = []
    while timer.date < timer.final_date:
        next(timer)
        nets = networks.active_networks(timer, policies)
        q = None
        if policies.quarantine_policies:
            q = policies.quarantine_policies.active_thresholds(timer)
        specs = []
        for net in nets:
            specs.append(O.NetSpec(name=net.name, edge_type=net.edge_type(), kind=net.kind,
                                   beta=net.beta_eff(policies, timer),
                                   prob=getattr(net, "leisure_probabilities", None)))
        steps.append(O.StepSpec(now=timer.now, dt=timer.duration, day_type=0 if timer.day_type == "weekday" else 1,
                                nets=specs, quarantine=q))
    return steps


def torch_noise(seed, n_calls, n_agents, device="cpu"):
    return [O.StepNoise(E=torch.from_numpy(E).to(device), u=torch.from_numpy(u).to(device),
                        z=torch.from_numpy(z).to(device)) for E, u, z in make_noise(seed, n_calls, n_agents)]


def profile_params(npz, device="cpu"):
    return {k: torch.from_numpy(npz["p_" + k]).to(device) for k in ("max_infectiousness", "shape", "rate", "shift")}


def make_leaf_networks(params):
    nets = InfectionNetworks.from_parameters(params)
    for key in nets.networks.keys():
        nets.networks[key].log_beta = torch.nn.Parameter(nets.networks[key].log_beta)
    return nets


RUNS = {"sample_default": 11, "sample_policies": 12, "sample_deadly": 13}
