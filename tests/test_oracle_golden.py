"""Pins oracle/gj_oracle.py against outputs of the REFERENCE's own code (tests/golden/*, generated
by make_golden.py): known-answer test, a single GradJune step and three full Runner trajectories
with gradients.  CPU only."""
import json

import numpy as np
import pytest
import torch

import helpers as H
from oracle import gj_oracle as O


def test_kat(golden_dir):
    """The reference's KAT (test/unit/infection_networks/test_base.py:21-44)."""
    k = json.load(open(golden_dir / "kat.json"))
    w = O.OracleWorld(n_agents=6, age=torch.zeros(6, dtype=torch.long), sex=torch.zeros(6, dtype=torch.long))
    w.edges["school"] = O.EdgeType(src=torch.tensor(k["src"]), dst=torch.tensor(k["dst"]),
                                   people=torch.tensor(k["people"]), n_groups=2)
    beta = 10.0 ** torch.tensor(float(k["log_beta"]))
    spec = O.StepSpec(now=0.0, dt=k["dt"], day_type=0, quarantine=None,
                      nets=[O.NetSpec("school", "school", O.KIND_PLAIN, beta)])
    q = O.not_infected_probs(w, spec, torch.tensor(k["transmission"]), torch.tensor(k["susceptibility"]),
                             torch.ones(6))
    assert np.array_equal(q.numpy(), np.array(k["q"], dtype=np.float32))          # bit-exact vs reference
    assert np.allclose(q.numpy(), np.array(k["expected_analytic"]))               # the reference's own assertion


def _step100_inputs(g, dtype=torch.float32):
    types = ["school", "company", "household"]
    w = H.oracle_world(g, types)
    state = {k: torch.from_numpy(g["pre_" + k]).to(dtype) for k in
             ("susceptibility", "is_infected", "infection_time", "current_stage", "next_stage", "time_to_next_stage")}
    params = {k: v.to(dtype) for k, v in H.profile_params(g).items()}
    return w, state, params


def test_step100(golden_dir):
    g = np.load(golden_dir / "step100.npz")
    w, state, params = _step100_inputs(g)
    import grad_june
    from grad_june.symptoms import SymptomsSampler
    sym = H.oracle_symptoms(SymptomsSampler.from_file())
    lb = {"household": torch.tensor(0.5, requires_grad=True), "company": torch.tensor(0.3, requires_grad=True),
          "school": torch.tensor(0.4, requires_grad=True)}
    kinds = {"household": O.KIND_HOUSEHOLD, "company": O.KIND_PLAIN, "school": O.KIND_PLAIN}
    nets = [O.NetSpec(n, n, kinds[n], 10.0 ** lb[n]) for n in g["order"]]
    # the default policies (social distancing from 2022-02-15) are inactive on 2022-02-04, quarantine list empty
    spec = O.StepSpec(now=float(g["now"]), dt=float(g["dt"]), day_type=0, nets=nets, quarantine=[])
    noise = H.torch_noise(8, 1, w.n_agents)[0]
    aux = {}
    O.step(w, state, params, spec, sym, noise, aux)
    assert np.array_equal(aux["transmission"].detach().numpy(), g["transmission"])
    assert np.array_equal(aux["q"].detach().numpy(), g["q"])
    for k in ("susceptibility", "is_infected", "infection_time", "current_stage", "next_stage", "time_to_next_stage"):
        assert np.array_equal(state[k].detach().numpy(), g["post_" + k]), k
    wl, w2 = torch.from_numpy(g["loss_w"]), torch.from_numpy(g["loss_w2"])
    loss = (state["is_infected"] * wl).sum() + (state["current_stage"] * w2).sum() \
        + 0.5 * (state["susceptibility"] * w2).sum() + 0.1 * (state["infection_time"] * wl).sum()
    loss.backward()
    grads = np.array([lb[k].grad.item() for k in ("household", "company", "school")])
    assert np.allclose(grads, g["grad_log_beta"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("tag", list(H.RUNS))
def test_runner_trajectory(golden_dir, tag):
    g = np.load(golden_dir / f"run_{tag}.npz")
    params, schedule = H.load_params(tag)
    arrays = np.load(golden_dir / "sample_world.npz")
    w = H.oracle_world(arrays, H.SAMPLE_TYPES)
    from grad_june.policies import Policies
    from grad_june.symptoms import SymptomsSampler
    nets = H.make_leaf_networks(params)
    policies = Policies.from_parameters(params)
    sym = H.oracle_symptoms(SymptomsSampler.from_parameters(params))
    steps = H.oracle_schedule(params, nets, policies)
    assert len(steps) == int(g["n_steps"])
    noises = H.torch_noise(H.RUNS[tag], len(steps) + 1, w.n_agents)
    log_frac = torch.tensor(float(params["infection_seed"]["log_fraction_initial_cases"]), requires_grad=True)
    trace = []
    res = O.run(w, H.profile_params(g), sym, log_frac, steps, noises,
                age_bins=params.get("age_bins_to_save", (0, 18, 65, 100)), trace=trace)
    assert np.array_equal(res["cases_per_timestep"].detach().numpy(), g["cases_per_timestep"])
    assert np.array_equal(res["deaths_per_timestep"].detach().numpy(), g["deaths_per_timestep"])
    assert np.array_equal(res["cases_by_age"].detach().numpy(), g["cases_by_age"])
    for key, gk in (("is_infected", "trace_is_infected"), ("current_stage", "trace_current_stage"),
                    ("next_stage", "trace_next_stage"), ("susceptibility", "trace_susceptibility")):
        mine = np.stack([t[key].numpy() for t in trace]).astype(np.uint8)
        assert np.array_equal(mine, g[gk]), key
    assert np.array_equal(trace[-1]["infection_time"].numpy(), g["final_infection_time"])
    assert np.array_equal(trace[-1]["time_to_next_stage"].numpy(), g["final_time_to_next_stage"])
    wc, wd, wa = g["loss_weights"]
    cba = res["cases_by_age"]
    loss = wc * res["cases_per_timestep"].sum() + wd * res["deaths_per_timestep"].sum() \
        + wa * (cba * torch.arange(1, cba.shape[1] + 1)).sum()
    loss.backward()
    assert np.isclose(loss.item(), float(g["loss"]), rtol=1e-6)
    names = [str(n) for n in g["net_names"]]
    grads = np.array([nets.networks[n].log_beta.grad.item() if nets.networks[n].log_beta.grad is not None else 0.0
                      for n in names])
    assert np.allclose(grads, g["grad_log_beta"], rtol=2e-5, atol=1e-30), (grads, g["grad_log_beta"])
    assert np.isclose(log_frac.grad.item(), float(g["grad_log_fraction"]), rtol=2e-5)
