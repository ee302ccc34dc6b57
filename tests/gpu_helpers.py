"""Builders shared by the GPU parity tests."""
import numpy as np
import torch

import helpers as H
from grad_june import GradJune, Runner, Timer
from grad_june.world import world_from_arrays


def make_runner(tag, device="cuda:0", renumber=False):
    """Runner on the reference's 769-agent sample world with the golden run's profile parameters.  ``renumber``:
    let Runner.get_data renumber the agents (household-contiguous inside their leisure cell); everything given or
    returned per agent in the LOADED numbering goes through layout_order_of / original_order."""
    from grad_june.world import layout_order_of
    g = np.load(H.GOLDEN / f"run_{tag}.npz")
    params, _ = H.load_params(tag)
    params["system"]["device"] = device
    params["system"]["renumber_agents"] = bool(renumber)
    arrays = np.load(H.GOLDEN / "sample_world.npz")
    data = Runner.get_data(params, data=world_from_arrays(arrays, H.SAMPLE_TYPES))
    assert ("original_index" in data["agent"]) == bool(renumber)
    data["agent"].infection_parameters = {k: layout_order_of(data, v.to(device)).contiguous()
                                          for k, v in H.profile_params(g).items()}
    model = GradJune.from_parameters(params)
    nets = model.infection_networks.networks
    for key in nets.keys():
        nets[key].log_beta = torch.nn.Parameter(nets[key].log_beta)
    log_frac = torch.nn.Parameter(torch.tensor(float(params["infection_seed"]["log_fraction_initial_cases"])))
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params),
                    log_fraction_initial_cases=log_frac, save_path=params["save_path"], parameters=params,
                    age_bins=params.get("age_bins_to_save", (0, 18, 65, 100)))
    return runner, g, params


def noise_provider(seed, n_calls, n_agents, device="cuda:0"):
    from noise import make_noise
    noise = [tuple(torch.from_numpy(x).to(device) for x in t) for t in make_noise(seed, n_calls, n_agents)]
    return lambda c: noise[c]


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else np.finfo(np.float64).tiny)
