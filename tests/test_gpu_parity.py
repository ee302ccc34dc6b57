"""GPU parity tests: the CUDA path (through the package API -> ctypes -> C ABI) against the golden
outputs of the reference and against oracle/gj_oracle.py on the same inputs and injected noise.

Tolerances (BASELINE.json north_star): infected masks / stage indices bit-exact; per-agent
probabilities and parameter gradients within 1e-5 relative (fp32).  A mask mismatch is tolerated
only as a *certified near-tie*: the two perturbed logits of that agent differ by less than a few
fp32 ulps, i.e. the draw is decided by the last bit of expf/logf (CPU vs GPU libm).
"""
import json

import numpy as np
import pytest
import torch

import helpers as H
from oracle import gj_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_library_loaded():
    from grad_june import _lib
    assert _lib.lib().gj_abi_version() == _lib.GJ_ABI_VERSION


def test_kat_through_module_api(golden_dir):
    """The reference's known-answer test, written against our classes like the reference writes it
    (test/unit/infection_networks/test_base.py:15-44)."""
    from grad_june import Timer
    from grad_june.infection_networks import InfectionNetworks, SchoolNetwork
    from grad_june.policies import Policies
    from grad_june.world import HeteroData, ToUndirected
    k = json.load(open(golden_dir / "kat.json"))
    networks = InfectionNetworks(school=SchoolNetwork(log_beta=np.log10(2.0)))
    data = HeteroData()
    data["agent"].id = torch.arange(6)
    data["agent"].age = torch.zeros(6, dtype=torch.long)
    data["agent"].sex = torch.zeros(6, dtype=torch.long)
    data["agent"].transmission = torch.tensor(k["transmission"])
    data["agent"].susceptibility = torch.tensor(k["susceptibility"])
    data["school"].id = torch.arange(2)
    data["school"].people = torch.tensor([2, 2])
    data["agent", "attends_school", "school"].edge_index = torch.vstack(
        (torch.arange(6), torch.tensor([0, 0, 0, 1, 1, 1])))
    data = ToUndirected()(data).to(DEV)
    timer = Timer(initial_day="2022-02-01", total_days=10, weekday_step_duration=(24,),
                  weekend_step_duration=(24,), weekday_activities=(("school",),), weekend_activities=(("school",),))
    q = networks(data=data, timer=timer, policies=Policies())
    expected = np.exp(-np.array([1.2, 2.4, 3.6, 1.5, 2.1, 3]))
    assert np.allclose(q.detach().cpu().numpy(), expected)
    assert np.allclose(q.detach().cpu().numpy(), np.array(k["q"], dtype=np.float32), rtol=2e-7, atol=0)


def _step100(golden_dir):
    from grad_june import GradJune, Timer
    from grad_june.infection_networks import CompanyNetwork, HouseholdNetwork, InfectionNetworks, SchoolNetwork
    from grad_june.world import world_from_arrays
    g = np.load(golden_dir / "step100.npz")
    data = world_from_arrays(g, ["school", "company", "household"])
    data["agent"].infection_parameters = H.profile_params(g)
    data["agent"].transmission = torch.zeros(100)
    for k in ("susceptibility", "is_infected", "infection_time"):
        data["agent"][k] = torch.from_numpy(g["pre_" + k])
    data["agent"].symptoms = {k: torch.from_numpy(g["pre_" + k]) for k in
                              ("current_stage", "next_stage", "time_to_next_stage")}
    data = data.to(DEV)
    nets = InfectionNetworks(
        household=HouseholdNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.5))),
        company=CompanyNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.3))),
        school=SchoolNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.4))))
    model = GradJune(infection_networks=nets, device=DEV)
    timer = Timer(initial_day="2022-02-01", total_days=10, weekday_step_duration=(24,),
                  weekday_activities=(("company", "school", "household"),))
    while timer.now < 3:
        next(timer)
    return g, data, model, nets, timer


def test_step100_forward_and_gradients(golden_dir):
    """One GradJune step on the 100-agent fixture world vs the reference's outputs."""
    from grad_june import ops
    from gpu_helpers import noise_provider
    g, data, model, nets, timer = _step100(golden_dir)
    with ops.inject_noise(noise_provider(8, 1, 100)):
        res = model(data=data, timer=timer)
    agent = res["agent"]
    T = agent.transmission.cpu().numpy()
    assert np.allclose(T, g["transmission"], rtol=RTOL, atol=1e-30)
    q = agent["not_infected_probs"].cpu().numpy()
    assert np.allclose(q, g["q"], rtol=RTOL, atol=0)
    for k in ("susceptibility", "is_infected"):
        assert np.array_equal(agent[k].detach().cpu().numpy(), g["post_" + k]), k
    for k in ("current_stage", "next_stage"):
        assert np.array_equal(agent.symptoms[k].detach().cpu().numpy(), g["post_" + k]), k
    assert np.allclose(agent.infection_time.detach().cpu().numpy(), g["post_infection_time"], rtol=1e-6)
    assert np.allclose(agent.symptoms["time_to_next_stage"].detach().cpu().numpy(), g["post_time_to_next_stage"],
                       rtol=1e-5)
    w = torch.from_numpy(g["loss_w"]).to(DEV)
    w2 = torch.from_numpy(g["loss_w2"]).to(DEV)
    loss = (agent.is_infected * w).sum() + (agent.symptoms["current_stage"] * w2).sum() \
        + 0.5 * (agent.susceptibility * w2).sum() + 0.1 * (agent.infection_time * w).sum()
    loss.backward()
    grads = np.array([nets[k].log_beta.grad.item() for k in ("household", "company", "school")])
    _, _, g64 = H.oracle_step100(torch.float64)
    H.assert_grad_parity(grads, g["grad_log_beta"], g64, rtol=RTOL, what="d/dlog_beta")


@pytest.mark.parametrize("tag", list(H.RUNS))
def test_runner_trajectory_vs_reference(tag):
    """Full Runner trajectories on the reference's 769-agent sample world (default config; policies
    with 8h/16h shifts; high-mortality variant) vs the reference's outputs, incl. d/dlog_beta (11) and
    d/dlog_fraction_initial_cases of a loss over cases, deaths and cases by age."""
    from grad_june import ops
    from gpu_helpers import make_runner, noise_provider
    runner, g, params = make_runner(tag, DEV)
    n_steps = int(g["n_steps"])
    with ops.inject_noise(noise_provider(H.RUNS[tag], n_steps + 1, runner.n_agents)):
        results, is_inf = runner()
    assert len(results["dates"]) == n_steps + 1
    assert np.array_equal(results["cases_per_timestep"].detach().cpu().numpy(), g["cases_per_timestep"])
    assert np.array_equal(results["deaths_per_timestep"].detach().cpu().numpy(), g["deaths_per_timestep"])
    bins = params.get("age_bins_to_save", (0, 18, 65, 100))
    cba = torch.stack([results[f"cases_by_age_{b:02d}"] for b in bins[1:]], dim=1)
    assert np.array_equal(cba.detach().cpu().numpy(), g["cases_by_age"])
    assert np.array_equal(is_inf.detach().cpu().numpy().astype(np.uint8), g["trace_is_infected"][-1])
    sym = runner.data["agent"].symptoms
    assert np.array_equal(sym["current_stage"].detach().cpu().numpy().astype(np.uint8), g["trace_current_stage"][-1])
    assert np.array_equal(sym["next_stage"].detach().cpu().numpy().astype(np.uint8), g["trace_next_stage"][-1])
    assert np.allclose(runner.data["agent"].infection_time.detach().cpu().numpy(), g["final_infection_time"], rtol=1e-6)
    assert np.allclose(sym["time_to_next_stage"].detach().cpu().numpy(), g["final_time_to_next_stage"], rtol=1e-5)
    wc, wd, wa = g["loss_weights"]
    loss = wc * results["cases_per_timestep"].sum() + wd * results["deaths_per_timestep"].sum() \
        + wa * (cba * torch.arange(1, cba.shape[1] + 1, device=DEV)).sum()
    loss.backward()
    assert np.isclose(loss.item(), float(g["loss"]), rtol=1e-6)
    nets = runner.model.infection_networks.networks
    names = [str(n) for n in g["net_names"]]
    grads = np.array([nets[n].log_beta.grad.item() if nets[n].log_beta.grad is not None else 0.0 for n in names])
    _, g64, gf64, same = H.oracle_run(tag, torch.float64)
    assert same, "fp64 witness left the golden trajectory"
    sens, sensf = H.run_sensitivity(tag)
    H.assert_grad_parity(grads, g["grad_log_beta"], g64, sens, rtol=RTOL, what="d/dlog_beta")
    H.assert_grad_parity(runner.log_fraction_initial_cases.grad.item(), float(g["grad_log_fraction"]), gf64, sensf,
                         rtol=RTOL, what="d/dlog_fraction_initial_cases")


def test_runner_is_deterministic():
    """Fixed-order segmented reductions: two runs with the same Philox key are bit-identical."""
    from grad_june import ops
    from gpu_helpers import make_runner
    runner, g, params = make_runner("sample_default", DEV)
    outs = []
    for _ in range(2):
        with ops.philox_seed(1234):
            results, is_inf = runner()
        loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum()
        for p in runner.model.infection_networks.networks.values():
            p.log_beta.grad = None
        loss.backward()
        grads = torch.stack([p.log_beta.grad for p in runner.model.infection_networks.networks.values()])
        outs.append((results["cases_per_timestep"].detach().clone(), is_inf.detach().clone(), grads.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert outs[0][0][-1] > outs[0][0][0]


def test_gradient_locality(golden_dir):
    """An infection that happened in a school must give d/dlog_beta_company == 0 exactly
    (reference test/unit/test_model.py:76-105)."""
    from grad_june import GradJune, Timer, ops
    from grad_june.infection import infect_people_at_indices
    from grad_june.infection_networks import CompanyNetwork, InfectionNetworks, SchoolNetwork
    from grad_june.world import HeteroData, ToUndirected
    torch.manual_seed(3)
    n = 100
    g = np.load(golden_dir / "step100.npz")
    data = HeteroData()
    data["agent"].id = torch.arange(n)
    data["agent"].age = torch.from_numpy(g["age"].astype(np.int64))
    data["agent"].sex = torch.from_numpy(g["sex"].astype(np.int64))
    data["agent"].infection_parameters = H.profile_params(g)
    data["agent"].transmission = torch.zeros(n)
    data["agent"].susceptibility = torch.ones(n)
    data["agent"].is_infected = torch.zeros(n)
    data["agent"].infection_time = torch.zeros(n)
    data["agent"].symptoms = {"current_stage": torch.ones(n, dtype=torch.long),
                              "next_stage": torch.ones(n, dtype=torch.long),
                              "time_to_next_stage": torch.zeros(n)}
    for name, lo in (("school", 0), ("company", 50)):
        data[name].id = torch.tensor([0])
        data[name].people = torch.tensor([50])
        data["agent", "attends_" + name, name].edge_index = torch.vstack(
            (torch.arange(lo, lo + 50), torch.zeros(50, dtype=torch.long)))
    data = ToUndirected()(data)
    data = infect_people_at_indices(data, list(range(0, 100, 10)))
    seeded = data["agent"].is_infected.clone()
    data = data.to(DEV)
    nets = InfectionNetworks(company=CompanyNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.5))),
                             school=SchoolNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.5))))
    model = GradJune(infection_networks=nets, device=DEV)
    timer = Timer(initial_day="2022-02-01", total_days=10, weekday_step_duration=(24,),
                  weekend_step_duration=(24,), weekday_activities=(("company", "school"),),
                  weekend_activities=(("company", "school"),))
    with ops.philox_seed(7):
        for _ in range(4):
            data = model(timer=timer, data=data)
            next(timer)
    cases = data["agent"]["is_infected"]
    new = ((cases.cpu() == 1.0) & (seeded == 0.0)).nonzero().flatten()
    in_school = [int(i) for i in new if i < 50]
    assert in_school, "nobody got infected at school"
    cases[in_school[0]].backward(retain_graph=True)
    assert nets["school"].log_beta.grad.item() != 0.0
    assert nets["company"].log_beta.grad.item() == 0.0


def test_user_defined_network_subclass(golden_dir):
    """A user subclass overriding the reference's masking hooks (base.py:47-59).  (1) One that restates the built-in
    behaviour must reproduce the fused path's probabilities; (2) a genuinely different mask (school susceptibility
    halved below age 10) must match a plain torch evaluation of base.py:61-84,118-141 and keep gradients flowing."""
    from grad_june import infection_networks as IN
    from grad_june import ops
    g, data, model, nets, timer = _step100(golden_dir)
    pre = {k: data["agent"][k].clone() for k in ("susceptibility", "is_infected", "infection_time")}
    pre_sym = {k: v.clone() for k, v in data["agent"].symptoms.items()}

    def reset():
        for k, v in pre.items():
            data["agent"][k] = v.clone()
        data["agent"].symptoms = {k: v.clone() for k, v in pre_sym.items()}

    with ops.philox_seed(3):
        model(data=data, timer=timer)
    q_fused = data["agent"]["not_infected_probs"].clone()
    assert model.kernel_family(data, timer) in ("throughput", "reference-order")

    class SchoolNetwork(IN.SchoolNetwork):            # (1) same behaviour, spelled out by the user
        def _get_transmissions(self, data, policies, timer):
            qp = policies.quarantine_policies
            return (qp.quarantine_mask if qp else 1.0) * data["agent"].transmission

    reset()
    nets.networks["school"] = SchoolNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.4)))
    assert nets.networks["school"].is_custom() and model.kernel_family(data, timer).startswith("modular")
    with ops.philox_seed(3):
        model(data=data, timer=timer)
    q_mod = data["agent"]["not_infected_probs"]
    assert torch.allclose(q_mod, q_fused, rtol=2e-6, atol=0)

    class SchoolNetwork(IN.SchoolNetwork):            # noqa: F811  (2) a different mask
        def _get_susceptibilities(self, data, policies, timer):
            young = (data["agent"].age < 10).float()
            return (1.0 - 0.5 * young) * data["agent"].susceptibility

    reset()
    nets.networks["school"] = SchoolNetwork(log_beta=torch.nn.Parameter(torch.tensor(0.4)))
    with ops.philox_seed(3):
        res = model(data=data, timer=timer)
    q_custom = res["agent"]["not_infected_probs"]
    # plain torch evaluation of the same step's pressure
    T = res["agent"].transmission.detach()
    s0, age = pre["susceptibility"], data["agent"].age
    lam = torch.zeros_like(s0)
    for name in timer.get_activity_order():
        ei = data["attends_" + name].edge_index
        people = data[name]["people"].float()
        pc = torch.clamp(1.0 / (people - 1), 0.0, 1.0)
        beta = 10.0 ** float(nets.networks[name].log_beta)
        sm = s0 * (1.0 - 0.5 * (age < 10).float()) if name == "school" else s0
        C = torch.zeros(len(people), device=DEV).index_add_(0, ei[1], T[ei[0]] * (beta * pc)[ei[1]])
        lam = lam + torch.zeros_like(s0).index_add_(0, ei[0], C[ei[1]] * sm[ei[0]])
    q_ref = torch.exp(-torch.clamp(lam, 1e-6, 100) * timer.duration)
    assert torch.allclose(q_custom, q_ref, rtol=1e-5, atol=0)
    assert not torch.allclose(q_custom, q_fused, rtol=1e-4)
    (res["agent"].is_infected.sum() + res["agent"].symptoms["current_stage"].sum()).backward()
    assert nets.networks["school"].log_beta.grad is not None and nets.networks["school"].log_beta.grad.item() != 0.0
    assert nets.networks["household"].log_beta.grad.item() != 0.0
