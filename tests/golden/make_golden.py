"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference package is imported from /root/reference (never copied) behind the
torch_geometric/h5py stand-ins of ``_pyg_shim.py``; every random draw on the hot path is replaced
by the deterministic noise of ``noise.py`` through torch-level monkey patches
(``Tensor.exponential_``, ``torch.bernoulli``, the dwell-time distributions' ``rsample``), so the
committed outputs are a pure function of (world, parameters, noise).  The tests replay the same
inputs through ``oracle/gj_oracle.py`` (CPU) and through the CUDA path (GPU).
"""
import copy
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import yaml

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(REF))

import _pyg_shim  # noqa: E402

_pyg_shim.install()
from noise import make_noise  # noqa: E402

import grad_june  # noqa: E402  (the reference)
from grad_june import GradJune, Runner, Timer  # noqa: E402
from grad_june.infection_networks import (  # noqa: E402
    CompanyNetwork, HouseholdNetwork, InfectionNetworks, SchoolNetwork)
from grad_june.paths import default_config_path  # noqa: E402
from grad_june.policies import Policies  # noqa: E402
from torch_geometric.data import HeteroData  # noqa: E402
import torch_geometric.transforms as T  # noqa: E402

assert "/root/reference" in grad_june.__file__, grad_june.__file__


# ----------------------------------------------------------------------------------
# noise injection
# ----------------------------------------------------------------------------------
class NoiseProvider:
    def __init__(self, noise):
        self.noise = noise
        self.c = -1

    def next_E(self):
        self.c += 1
        return torch.from_numpy(self.noise[self.c][0])

    def u(self):
        return torch.from_numpy(self.noise[self.c][1])

    def z(self, row):
        return torch.from_numpy(self.noise[self.c][2][row])


class inject:
    """Context manager patching torch's samplers with the provider's arrays."""

    def __init__(self, provider, symptoms_sampler):
        self.p = provider
        self.ss = symptoms_sampler

    def __enter__(self):
        p = self.p
        self._exp = torch.Tensor.exponential_
        self._bern = torch.bernoulli

        def exponential_(t, lambd=1.0, generator=None):
            e = p.next_E()
            assert e.shape == t.shape, (e.shape, t.shape)
            t.copy_(e)
            return t

        def bernoulli(probs, *a, **k):
            return (p.u() < probs).to(probs.dtype)

        torch.Tensor.exponential_ = exponential_
        torch.bernoulli = bernoulli
        self._dists = []
        for kind, table in ((0, self.ss.stage_transition_times), (1, self.ss.recovery_times)):
            for i, d in table.items():
                if d is None:
                    continue
                row = (i - 2) * 2 + kind

                def rsample(shape, d=d, row=row):
                    x = d.base_dist.loc + p.z(row) * d.base_dist.scale if hasattr(d, "base_dist") \
                        else d.loc + p.z(row) * d.scale
                    return torch.exp(x) if hasattr(d, "base_dist") else x

                self._dists.append(d)
                d.rsample = rsample
        return self

    def __exit__(self, *a):
        torch.Tensor.exponential_ = self._exp
        torch.bernoulli = self._bern
        for d in self._dists:
            del d.rsample


# ----------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------
def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def world_arrays(data, types):
    out = {"age": data["agent"].age.numpy().astype(np.int16), "sex": data["agent"].sex.numpy().astype(np.int8)}
    for t in types:
        ei = data["attends_" + t].edge_index.numpy()
        rev = data["rev_attends_" + t].edge_index.numpy()
        assert (rev[0] == ei[1]).all() and (rev[1] == ei[0]).all()   # same order, rows swapped
        out[f"{t}_src"] = ei[0].astype(np.int32)
        out[f"{t}_dst"] = ei[1].astype(np.int32)
        out[f"{t}_people"] = np.asarray(data[t]["people"]).astype(np.int32)
        out[f"{t}_ngroups"] = np.int32(len(data[t]["id"]))
    return out


def record_schedule(runner_or_model, timer, policies, networks):
    order = timer.get_activity_order()
    if policies.close_venue_policies:
        order = policies.close_venue_policies.apply(edge_types=order, timer=timer)
    betas = {}
    for name in order:
        net = networks[name]
        beta = 10.0 ** net.log_beta
        if policies.interaction_policies:
            beta = policies.interaction_policies.apply(beta=beta, name=net.name, timer=timer)
        betas[name] = float(beta)
    q = None
    if policies.quarantine_policies:
        q = [float(p.stage_threshold) for p in policies.quarantine_policies.policies
             if p.is_active(timer.date)]
    return {"now": timer.now, "dt": timer.duration, "day_type": timer.day_type,
            "date": timer.date.isoformat(), "order": list(order), "beta": betas, "quarantine": q}


STATE_KEYS = ("susceptibility", "is_infected", "infection_time")
SYM_KEYS = ("current_stage", "next_stage", "time_to_next_stage")


def snapshot(data):
    out = {k: np32(data["agent"][k]) for k in STATE_KEYS}
    out.update({k: np32(data["agent"].symptoms[k]) for k in SYM_KEYS})
    return out


# ----------------------------------------------------------------------------------
# 1. the reference's known-answer test (test/unit/infection_networks/test_base.py:21-44)
# ----------------------------------------------------------------------------------
def gen_kat():
    sn = SchoolNetwork(log_beta=np.log10(2.0))
    nets = InfectionNetworks(school=sn)
    data = HeteroData()
    data["agent"].id = torch.arange(6)
    data["agent"].transmission = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.6])
    data["agent"].susceptibility = torch.tensor([1, 2, 3, 0.5, 0.7, 1.0])
    data["school"].id = torch.arange(2)
    data["school"].people = torch.tensor([2, 2])
    data["agent", "attends_school", "school"].edge_index = torch.vstack(
        (torch.arange(6), torch.tensor([0, 0, 0, 1, 1, 1])))
    data = T.ToUndirected()(data)
    timer = Timer(initial_day="2022-02-01", total_days=10, weekday_step_duration=(24,),
                  weekend_step_duration=(24,), weekday_activities=(("school",),),
                  weekend_activities=(("school",),))
    q = nets(data=data, timer=timer, policies=Policies())
    expected = np.exp(-np.array([1.2, 2.4, 3.6, 1.5, 2.1, 3]))
    assert np.allclose(q.detach().numpy(), expected)
    return {"q": np32(q).tolist(), "expected_analytic": expected.tolist(),
            "transmission": [0.1, 0.2, 0.3, 0.4, 0.5, 0.6], "susceptibility": [1, 2, 3, 0.5, 0.7, 1.0],
            "people": [2, 2], "src": list(range(6)), "dst": [0, 0, 0, 1, 1, 1],
            "log_beta": float(np.log10(2.0)), "dt": 1.0}


# ----------------------------------------------------------------------------------
# 2. one GradJune step on a conftest-style 100-agent world (test/conftest.py:36-89,
#    test/unit/test_model.py:15-53), with gradients wrt the three log_betas
# ----------------------------------------------------------------------------------
def gen_step100(seed=7):
    rng = np.random.default_rng(seed)
    n = 100
    torch.manual_seed(999)
    sampler = grad_june.TransmissionSampler.from_file()
    data = HeteroData()
    data["agent"].id = torch.arange(0, n)
    data["agent"].age = torch.from_numpy(rng.integers(0, 100, n))
    data["agent"].sex = torch.from_numpy(rng.integers(0, 2, n))
    vals = sampler(n)
    inf_params = {k: vals[i] for i, k in enumerate(("max_infectiousness", "shape", "rate", "shift"))}
    data["agent"].infection_parameters = inf_params
    data["agent"].transmission = torch.zeros(n)
    data["agent"].susceptibility = torch.ones(n)
    data["agent"].is_infected = torch.zeros(n)
    data["agent"].infection_time = torch.zeros(n)
    data["agent"].symptoms = {"current_stage": torch.ones(n, dtype=torch.long),
                              "next_stage": torch.ones(n, dtype=torch.long),
                              "time_to_next_stage": torch.zeros(n)}
    # shuffled (unsorted) edge lists, uneven groups; people deliberately != member count for school
    def edges(n_groups):
        dst = rng.integers(0, n_groups, n)
        perm = rng.permutation(n)
        return torch.from_numpy(np.stack([np.arange(n)[perm], dst[perm]]))
    for name, ng in (("school", 4), ("company", 7), ("household", 25)):
        data[name].id = torch.arange(ng)
        ei = edges(ng)
        data["agent", "attends_" + name, name].edge_index = ei
        cnt = np.bincount(ei[1].numpy(), minlength=ng)
        data[name].people = torch.from_numpy(cnt if name != "school" else cnt + 3)
    data = T.ToUndirected()(data)
    from grad_june.infection import infect_people_at_indices
    data = infect_people_at_indices(data, list(range(0, 100, 10)))
    # a few agents already in later stages / recovered, to exercise the symptoms machine
    cur = data["agent"].symptoms["current_stage"].clone()
    nxt = data["agent"].symptoms["next_stage"].clone()
    ttn = data["agent"].symptoms["time_to_next_stage"].clone()
    for k, (c, nx, tt) in enumerate([(3, 4, 1.0), (4, 5, 2.5), (5, 6, 0.2), (6, 7, 0.1), (7, 7, 0.0),
                                     (0, 0, 0.0), (2, 3, 5.0), (3, 0, 0.5)]):
        a = 5 + 10 * k
        cur[a], nxt[a], ttn[a] = c, nx, tt
        data["agent"].is_infected[a] = 1.0
        data["agent"].susceptibility[a] = 0.0
        data["agent"].infection_time[a] = -1.5
    data["agent"].symptoms = {"current_stage": cur, "next_stage": nxt, "time_to_next_stage": ttn}

    nets = InfectionNetworks(
        household=HouseholdNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.5))),
        company=CompanyNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.3))),
        school=SchoolNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.4))))
    model = GradJune(infection_networks=nets)
    timer = Timer(initial_day="2022-02-01", total_days=10, weekday_step_duration=(24,),
                  weekday_activities=(("company", "school", "household"),))
    while timer.now < 3:
        next(timer)
    out = world_arrays(data, ["school", "company", "household"])
    out.update({"p_" + k: np32(v) for k, v in inf_params.items()})
    out.update({"pre_" + k: v for k, v in snapshot(data).items()})
    out["now"], out["dt"] = np.float64(timer.now), np.float64(timer.duration)
    out["order"] = np.array(timer.get_activity_order())
    noise = make_noise(seed + 1, 1, n)
    prov = NoiseProvider(noise)
    captured = {}
    orig = model.infection_networks.forward

    def nets_fwd(*a, **k):
        q = orig(*a, **k)
        captured["q"] = q
        return q
    model.infection_networks.forward = nets_fwd
    with inject(prov, model.symptoms_updater.symptoms_sampler):
        res = model(data=data, timer=timer)
    out["transmission"] = np32(res["agent"].transmission)
    out["q"] = np32(captured["q"])
    out.update({"post_" + k: v for k, v in snapshot(res).items()})
    w = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
    w2 = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
    loss = (res["agent"].is_infected * w).sum() + (res["agent"].symptoms["current_stage"] * w2).sum() \
        + 0.5 * (res["agent"].susceptibility * w2).sum() + 0.1 * (res["agent"].infection_time * w).sum()
    loss.backward()
    out["loss_w"], out["loss_w2"] = w.numpy(), w2.numpy()
    out["loss"] = np.float32(loss.item())
    out["grad_log_beta"] = np.array([nets[k].log_beta.grad.item() for k in ("household", "company", "school")],
                                    dtype=np.float64)
    np.savez_compressed(HERE / "step100.npz", **out)
    print("step100: new infected", int((out["post_is_infected"] - out["pre_is_infected"]).sum()),
          "grads", out["grad_log_beta"])


# ----------------------------------------------------------------------------------
# 3./4. full Runner trajectories on the reference's sample world (test/data/data.pkl)
# ----------------------------------------------------------------------------------
POLICY_BLOCK = {
    "interaction": {"social_distancing": {
        1: {"start_date": "2022-02-05", "end_date": "2022-02-12",
            "beta_factors": {"school": 0.5, "company": 0.4, "pub": 0.3, "all": 0.8}},
        2: {"start_date": "2022-02-08", "end_date": "2022-02-20", "beta_factors": {"household": 1.5, "visit": 0.1}}}},
    "quarantine": {"quarantine": {
        1: {"start_date": "2022-02-04", "end_date": "2022-02-25", "stage_threshold": 4}}},
    "close_venue": {"close_venue": {
        1: {"start_date": "2022-02-06", "end_date": "2022-02-10", "names": ["school", "pub", "cinema", "gym"]}}},
}


def gen_run(tag, params, noise_seed, loss_weights):
    torch.manual_seed(999)
    np.random.seed(999)
    runner = Runner.from_parameters(params)
    nets = runner.model.infection_networks.networks
    names = list(nets.keys())
    for key in names:
        nets[key].log_beta = torch.nn.Parameter(nets[key].log_beta)
    runner.log_fraction_initial_cases = torch.nn.Parameter(torch.tensor(float(runner.log_fraction_initial_cases)))
    n = runner.n_agents
    # count timesteps
    t = Timer.from_parameters(params)
    n_steps = 0
    while t.date < t.final_date:
        next(t)
        n_steps += 1
    noise = make_noise(noise_seed, n_steps + 1, n)
    prov = NoiseProvider(noise)
    schedule, trace = [], []
    orig_forward = runner.model.forward

    def fwd(data, timer):
        schedule.append(record_schedule(runner, timer, runner.model.policies, runner.model.infection_networks))
        out = orig_forward(data, timer)
        trace.append(snapshot(out))
        return out
    runner.model.forward = fwd
    orig_seed = runner.set_initial_cases

    def seed_and_trace():
        orig_seed()
        trace.append(snapshot(runner.data))
    runner.set_initial_cases = seed_and_trace
    with inject(prov, runner.model.symptoms_updater.symptoms_sampler):
        results, is_inf = runner()
    assert prov.c == n_steps, (prov.c, n_steps)
    wc, wd, wa = loss_weights
    bins = params.get("age_bins_to_save", (0, 18, 65, 100))
    cba = torch.stack([results[f"cases_by_age_{b:02d}"] for b in bins[1:]], dim=1)
    loss = wc * results["cases_per_timestep"].sum() + wd * results["deaths_per_timestep"].sum() \
        + wa * (cba * torch.arange(1, cba.shape[1] + 1)).sum()
    loss.backward()
    out = {
        "n_steps": np.int32(n_steps),
        "cases_per_timestep": np32(results["cases_per_timestep"]),
        "deaths_per_timestep": np32(results["deaths_per_timestep"]),
        "cases_by_age": np32(cba),
        "loss": np.float64(loss.item()),
        "loss_weights": np.array(loss_weights, dtype=np.float64),
        "net_names": np.array(names),
        "log_beta": np.array([float(nets[k].log_beta) for k in names]),
        "grad_log_beta": np.array([nets[k].log_beta.grad.item() for k in names], dtype=np.float64),
        "grad_log_fraction": np.float64(runner.log_fraction_initial_cases.grad.item()),
        "trace_is_infected": np.stack([s["is_infected"] for s in trace]).astype(np.uint8),
        "trace_current_stage": np.stack([s["current_stage"] for s in trace]).astype(np.uint8),
        "trace_next_stage": np.stack([s["next_stage"] for s in trace]).astype(np.uint8),
        "trace_susceptibility": np.stack([s["susceptibility"] for s in trace]).astype(np.uint8),
        "final_infection_time": trace[-1]["infection_time"],
        "final_time_to_next_stage": trace[-1]["time_to_next_stage"],
    }
    for k in ("max_infectiousness", "shape", "rate", "shift"):
        out["p_" + k] = np32(runner.data["agent"].infection_parameters[k])
    # is_infected may exceed 1 (re-infection through the 1e-6 floor) but stays a small integer
    assert np.stack([s["is_infected"] for s in trace]).max() < 255
    np.savez_compressed(HERE / f"run_{tag}.npz", **out)
    with open(HERE / f"schedule_{tag}.json", "w") as f:
        json.dump({"params": params, "schedule": schedule}, f, indent=1, default=str)
    print(f"run_{tag}: steps={n_steps} cases={out['cases_per_timestep'][[0, -1]]} "
          f"deaths={out['deaths_per_timestep'][-1]} grad={out['grad_log_beta']} gfrac={out['grad_log_fraction']}")
    return runner


def main():
    os.chdir(HERE)
    with open(HERE / "kat.json", "w") as f:
        json.dump(gen_kat(), f, indent=1)
    gen_step100()

    with open(default_config_path) as f:
        params = yaml.safe_load(f)
    with open(HERE / "reference_default_params.json", "w") as f:
        json.dump(params, f, indent=1, default=str, sort_keys=True)

    # sample world, converted once to plain arrays (the pickle itself needs torch_geometric to load)
    runner = gen_run("sample_default", copy.deepcopy(params), 11, (1.0, 3.0, 0.25))
    types = ["household", "company", "school", "university", "care_home", "leisure"]
    arrays = world_arrays(runner.data, types)
    arrays["ethnicity"] = np.asarray(runner.data["agent"].ethnicity)
    np.savez_compressed(HERE / "sample_world.npz", **arrays)

    p2 = copy.deepcopy(params)
    p2["policies"] = POLICY_BLOCK
    p2["timer"]["total_days"] = 20
    p2["timer"]["step_duration"] = {"weekday": {0: 8, 1: 16}, "weekend": {0: 24}}
    wk0 = ["company", "school", "university", "care_home", "household"]
    wk1 = ["pub", "grocery", "gym", "cinema", "visit", "care_visit", "care_home", "household"]
    p2["timer"]["step_activities"] = {"weekday": {0: wk0, 1: wk1}, "weekend": {0: wk1}}
    p2["infection_seed"]["log_fraction_initial_cases"] = -1.3
    gen_run("sample_policies", p2, 12, (0.5, 2.0, 0.1))

    # high-mortality, fast-progression variant so that deaths (and their gradient) are exercised,
    # with one Normal dwell-time distribution (test/unit/test_symptoms.py uses Normal too)
    p3 = copy.deepcopy(params)
    p3["timer"]["total_days"] = 25
    p3["policies"] = {}
    sy = p3["symptoms"]
    for st, pr in (("infectious", 0.9), ("symptomatic", 0.8), ("severe", 0.85), ("critical", 0.7)):
        sy["stage_transition_probabilities"][st] = {"0-50": pr, "50-100": min(1.0, pr + 0.1)}
    for st, (loc, sc) in (("exposed", (0.3, 0.3)), ("infectious", (0.1, 0.4)), ("symptomatic", (0.5, 0.3)),
                          ("critical", (0.4, 0.5))):
        sy["stage_transition_times"][st] = {"dist": "LogNormal", "loc": loc, "scale": sc}
    sy["stage_transition_times"]["severe"] = {"dist": "Normal", "loc": 1.5, "scale": 0.2}
    p3["infection_seed"]["log_fraction_initial_cases"] = -0.8
    gen_run("sample_deadly", p3, 13, (0.2, 5.0, 0.3))


if __name__ == "__main__":
    main()
