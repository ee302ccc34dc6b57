"""GPU parity at larger sizes: a synthetic England-like world (all six edge types, eleven networks,
quarantine active) stepped once from a mid-epidemic state, teacher-forced against the oracle with the
kernel's own Philox noise (gj_philox_fill), plus size-independent properties."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import gj_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _params(device=DEV):
    from grad_june.default_config import default_parameters
    p = default_parameters()
    p["system"]["device"] = device
    p["policies"] = {"quarantine": {"quarantine": {1: {"start_date": "2022-01-01", "end_date": "2023-01-01",
                                                       "stage_threshold": 4}}}}
    return p


def _mid_epidemic_state(n, now, seed, device):
    return H.mid_epidemic_state(n, now, seed, device)


def _irregular(data):
    """Make the synthetic world exercise the rare layout paths: households of up to 20 members (the range tier's
    loop beyond its eight unrolled neighbours) and agents with two or three generic-tier edges (the agent-major
    CSR walk behind the one-entry-per-agent word)."""
    n = len(data["agent"].id)
    dev = data["agent"].age.device
    g = torch.Generator(device="cpu").manual_seed(123)
    sizes = torch.randint(1, 21, (n // 2 + 8,), generator=g).to(dev)
    ends = torch.cumsum(sizes, 0)
    n_hh = int(torch.searchsorted(ends, torch.tensor([n], device=dev))[0]) + 1
    ids = torch.arange(n, device=dev)
    hh = torch.searchsorted(ends[:n_hh].contiguous(), ids, right=True)
    data["household"].id = torch.arange(n_hh, device=dev)
    data["household"].people = torch.bincount(hh, minlength=n_hh)
    data["attends_household"].edge_index = torch.stack((ids, hh))
    for src_t, dst_t, frac in (("company", "university", 0.10), ("school", "care_home", 0.05), ("company", "care_home", 0.03)):
        members = data["attends_" + src_t].edge_index[0]
        pick = members[torch.rand(members.numel(), generator=g).to(dev) < frac]
        G = len(data[dst_t].id)
        grp = torch.randint(0, G, (pick.numel(),), generator=g).to(dev)
        ei = torch.cat((data["attends_" + dst_t].edge_index, torch.stack((pick, grp))), dim=1)
        data["attends_" + dst_t].edge_index = ei
        data[dst_t].people = torch.bincount(ei[1], minlength=G)
    return data


def _setup(n_agents, seed=1, mutate=None):
    from grad_june import GradJune, Timer
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    params = _params()
    torch.manual_seed(seed)
    data = make_synthetic_world(n_agents, seed=seed, device=DEV)
    if mutate is not None:
        data = mutate(data)
    data = Runner.get_data(params, data=data)
    model = GradJune.from_parameters(params)
    for net in model.infection_networks.networks.values():
        net.log_beta = torch.nn.Parameter(net.log_beta)
    timer = Timer.from_parameters(params)
    for _ in range(8):
        next(timer)
    state = _mid_epidemic_state(n_agents, timer.now, seed + 5, DEV)
    for k in ("susceptibility", "is_infected", "infection_time"):
        data["agent"][k] = state[k]
    data["agent"].symptoms = {k: state[k] for k in ("current_stage", "next_stage", "time_to_next_stage")}
    return params, data, model, timer, state


def _oracle_inputs(params, data, model, timer, state, device):
    return H.oracle_step_inputs(params, data, model, timer, state, device)


@pytest.mark.parametrize("n_agents,oracle_device", [(30_000, "cpu"), (400_000, DEV)])
def test_teacher_forced_step_vs_oracle(n_agents, oracle_device):
    from grad_june import ops
    params, data, model, timer, state = _setup(n_agents)
    seed = 2024
    with ops.philox_seed(seed):
        res = model(data=data, timer=timer)
    agent = res["agent"]
    E, u, z = ops.philox_fill(seed, 0, n_agents, DEV)
    w, nets, spec, sym, prof, st = _oracle_inputs(params, data, model, timer, state, oracle_device)
    noise = O.StepNoise(E=E.to(oracle_device), u=u.to(oracle_device), z=z.to(oracle_device).expand(10, n_agents))
    aux = {}
    O.step(w, st, prof, spec, sym, noise, aux)

    def cpu(t):
        return t.detach().cpu().numpy()

    T, To = cpu(agent.transmission), cpu(aux["transmission"])
    assert np.allclose(T, To, rtol=1e-5, atol=1e-30)
    q, qo = cpu(agent["not_infected_probs"]), cpu(aux["q"])
    assert np.max(np.abs(q - qo) / qo) <= 1e-5, np.max(np.abs(q - qo) / qo)
    n, no = cpu(agent["new_infected"]), cpu(aux["new_infected"])
    mism = np.nonzero(n != no)[0]
    # every mask mismatch must be a certified near-tie of the two perturbed logits (helpers.certify_near_ties)
    worst = H.certify_near_ties(qo, E.cpu().numpy(), mism, what=f"teacher-forced step, {n_agents} agents")
    assert len(mism) <= max(2, int(1e-6 * n_agents)), len(mism)
    H.report(f"teacher-forced step {n_agents} agents: mask mismatches / worst gap (units of the certificate)",
             {"mismatches": int(len(mism)), "worst": worst, "q_max_rel": float(np.max(np.abs(q - qo) / qo))})
    ok = np.ones(n_agents, dtype=bool)
    ok[mism] = False
    assert n.sum() > 0.001 * n_agents
    for mine, theirs, exact in ((agent.susceptibility, st["susceptibility"], True),
                                (agent.is_infected, st["is_infected"], True),
                                (agent.symptoms["current_stage"], st["current_stage"], True),
                                (agent.symptoms["next_stage"], st["next_stage"], True),
                                (agent.infection_time, st["infection_time"], False),
                                (agent.symptoms["time_to_next_stage"], st["time_to_next_stage"], False)):
        a, b = cpu(mine)[ok], cpu(theirs)[ok]
        if exact:
            assert np.array_equal(a, b)
        else:
            assert np.allclose(a, b, rtol=2e-6, atol=1e-6)
    # stages really moved (the symptoms machine was exercised)
    assert (cpu(agent.symptoms["current_stage"]) != cpu(state["current_stage"])).sum() > 0.01 * n_agents

    # ---- gradients, unconditionally: an agent whose draw differs (certified near-tie above) is left out of the loss
    #      on both sides, so the comparison never depends on whether a flip happened -------------------------------
    gw = torch.Generator(device="cpu").manual_seed(5)
    keep = torch.from_numpy(ok.astype(np.float32))
    wts = [torch.rand(n_agents, generator=gw) * keep for _ in range(4)]
    keys = list(model.infection_networks.networks.keys())

    def loss_of(inf, cur, s, tinf, dev, dtype=torch.float32):
        w0, w1, w2, w3 = (t.to(device=dev, dtype=dtype) for t in wts)
        return (inf * w0).sum() + (cur * w1).sum() + 0.5 * (s * w2).sum() + 0.05 * (tinf * w3).sum()

    def oracle_grads(dtype, perturb=None):
        wo, netso, speco, symo, profo, sto = _oracle_inputs(params, data, model, timer, state, oracle_device)
        for sp in speco.nets:
            sp.beta = sp.beta.to(dtype)
        sto = {k: v.to(dtype) for k, v in sto.items()}
        profo = {k: v.to(dtype) for k, v in profo.items()}
        nz = O.StepNoise(E=noise.E.to(dtype), u=noise.u.to(dtype), z=noise.z.to(dtype))
        if perturb is None:
            O.step(wo, sto, profo, speco, symo, nz)
        else:
            with H.perturb_q(perturb):
                O.step(wo, sto, profo, speco, symo, nz)
        loss_of(sto["is_infected"], sto["current_stage"], sto["susceptibility"], sto["infection_time"],
                oracle_device, dtype).backward()
        gr = np.array([netso.networks[k].log_beta.grad.item() if netso.networks[k].log_beta.grad is not None
                       else 0.0 for k in keys])
        return gr, torch.equal((sto["is_infected"].float() * keep.to(oracle_device)),
                               (st["is_infected"] * keep.to(oracle_device)))

    loss_of(agent.is_infected, agent.symptoms["current_stage"], agent.susceptibility, agent.infection_time, DEV).backward()
    mine = np.array([model.infection_networks.networks[k].log_beta.grad.item() for k in keys])
    ref, _ = oracle_grads(torch.float32)
    g64, same = oracle_grads(torch.float64)
    assert same, "fp64 witness drew different masks"
    sens = np.zeros_like(ref)
    for sd in (1, 2, 3):
        gp, _ = oracle_grads(torch.float32, perturb=sd)
        sens = np.maximum(sens, np.abs(gp - ref))
    H.assert_grad_parity(mine, ref, g64, sens, rtol=1e-5, what=f"teacher-forced step {n_agents} agents: d/dlog_beta")


def test_properties_at_scale():
    """Size-independent checks on a 2M-agent world: reductions equal recomputed sums, run-to-run
    bit-reproducibility, and pressure is linear in beta (q(2*beta) == q(beta)**2 where unclamped)."""
    from grad_june import ops
    n_agents = 2_000_000
    params, data, model, timer, state = _setup(n_agents, seed=3)

    def reset():
        for k in ("susceptibility", "is_infected", "infection_time"):
            data["agent"][k] = state[k]
        data["agent"].symptoms = {k: state[k] for k in ("current_stage", "next_stage", "time_to_next_stage")}

    outs = []
    for rep in range(2):
        reset()
        with ops.philox_seed(99):
            _, red = model.step(data, timer, age_bins=(0, 18, 65, 100))
        outs.append((data["agent"].is_infected.clone(), data["agent"]["not_infected_probs"].clone(), red.clone(),
                     data["agent"].symptoms["current_stage"].clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    inf, q, red, cur = outs[0]
    age = data["agent"].age
    assert red[0].item() == pytest.approx(inf.double().sum().item(), rel=1e-7)
    assert red[1].item() == pytest.approx((cur == 7).double().sum().item(), rel=1e-7)
    for i, (lo, hi) in enumerate(((0, 18), (18, 65), (65, 100))):
        assert red[2 + i].item() == pytest.approx((inf.double() * ((age > lo) & (age < hi))).sum().item(), rel=1e-7)
    assert float(q.min()) > 0 and float(q.max()) <= 1.0
    # linearity in beta
    reset()
    with torch.no_grad():
        for net in model.infection_networks.networks.values():
            net.log_beta += np.log10(2.0)
    with ops.philox_seed(99):
        model.step(data, timer)
    q2 = data["agent"]["not_infected_probs"]
    mid = (q > 1e-3) & (q < 0.99)       # away from the 1e-6 floor and the 100 cap
    assert int(mid.sum()) > 1000
    assert torch.allclose(q2[mid], q[mid] ** 2, rtol=2e-5)


@pytest.mark.parametrize("irregular,pipelined", [(False, True), (True, True), (True, False)])
def test_throughput_mode_matches_reference_order(irregular, pipelined):
    """The throughput-mode kernels (re-associated sums, hardware log2/exp2 draw; used with in-kernel Philox
    noise), pipelined and register-batched, against the reference-order kernels on the same state and the same
    Philox stream — on the regular synthetic world and on one with 20-member households and multi-edge agents."""
    from grad_june import _lib, ops
    from grad_june.world import TIER_RANGE, get_device_world
    n_agents = 300_000
    params, data, model, timer, state = _setup(n_agents, seed=11, mutate=_irregular if irregular else None)
    if irregular:
        world = get_device_world(data, DEV)
        assert world.type_tier[world.types.index("household")] == TIER_RANGE
        assert int((world.ent1.long() & 0xFFFFFFFF == 0xFFFFFFFE).sum()) > 1000      # agents with several generic edges
        assert int(data["household"].people.max()) > 8
    prev = _lib.pipeline_enable(None)
    _lib.pipeline_enable(pipelined, lookahead=False)
    try:
        _compare_throughput_with_reference_order(n_agents, data, model, timer, state)
    finally:
        _lib.pipeline_enable(*prev)


def _compare_throughput_with_reference_order(n_agents, data, model, timer, state):
    from grad_june import ops
    outs, graphs = {}, {}
    for mode in ("exact", "fast"):
        for k in ("susceptibility", "is_infected", "infection_time"):
            data["agent"][k] = state[k]
        data["agent"].symptoms = {k: state[k] for k in ("current_stage", "next_stage", "time_to_next_stage")}
        ops.EXACT_ORDER = mode == "exact"
        try:
            with ops.philox_seed(77):
                _, red = model.step(data, timer, age_bins=(0, 18, 65, 100))
        finally:
            ops.EXACT_ORDER = False
        agent = data["agent"]
        graphs[mode] = (agent.is_infected, agent.susceptibility, agent.infection_time, agent.symptoms["current_stage"])
        outs[mode] = dict(q=agent["not_infected_probs"].cpu().numpy(), n=agent["new_infected"].cpu().numpy(),
                          cur=agent.symptoms["current_stage"].detach().cpu().numpy(),
                          ttn=agent.symptoms["time_to_next_stage"].detach().cpu().numpy(), red=red.detach().cpu().numpy())
    e, f = outs["exact"], outs["fast"]
    assert np.max(np.abs(e["q"] - f["q"]) / e["q"]) < 2e-6
    mism = np.nonzero(e["n"] != f["n"])[0]
    assert len(mism) <= 3, len(mism)        # only near-ties may differ
    ok = np.ones(n_agents, dtype=bool)
    ok[mism] = False
    assert np.array_equal(e["cur"][ok], f["cur"][ok])
    assert np.allclose(e["ttn"][ok], f["ttn"][ok], rtol=1e-6, atol=1e-6)
    assert np.allclose(e["red"], f["red"], atol=len(mism) + 0.5)
    # gradients, unconditionally: per-agent losses with the flipped agents (if any) left out on both sides
    keep = torch.from_numpy(ok.astype(np.float32)).to(DEV)
    ramp = torch.linspace(0, 1, n_agents, device=DEV) * keep
    grads = {}
    for mode in ("exact", "fast"):
        for net in model.infection_networks.networks.values():
            net.log_beta.grad = None
        inf, s, tinf, cur = graphs[mode]
        loss = (inf * keep).sum() + (s * ramp).sum() + 0.01 * (tinf * keep).sum() + (cur * keep).sum() \
            + 0.3 * (inf * keep * (data["agent"].age < 18)).sum()
        loss.backward()
        grads[mode] = torch.stack([net.log_beta.grad for net in model.infection_networks.networks.values()]).cpu().numpy()
    assert np.abs(grads["exact"]).max() > 0
    assert np.allclose(grads["exact"], grads["fast"], rtol=2e-4, atol=1e-7 * np.abs(grads["exact"]).max())
    H.report("throughput vs reference-order kernels, one step: mask mismatches", int(len(mism)))


@pytest.mark.parametrize("quarantine", [False, True])
def test_compacted_transmission_pass(quarantine):
    """k_lean_transmission_c (the infectious agents compacted per warp: the default) against k_lean_transmission on the
    same mid-epidemic state and Philox stream: transmissions bit-identical; the per-thread partial sums of the leisure
    channels are associated differently, so q agrees to fp32 rounding and draws differ only at near-ties."""
    from grad_june import _lib, ops
    n_agents = 250_003
    params, data, model, timer, state = _setup(n_agents, seed=21)
    if not quarantine:      # _setup's parameters carry an active quarantine policy
        from grad_june.policies import Policies
        params = dict(params)
        params["policies"] = {}
        model.policies = Policies.from_parameters(params)
    assert model.kernel_family(data, timer) == "throughput"
    outs = {}
    prev = _lib.pipeline_enable(None)
    try:
        for uncompacted in (True, False):
            _lib.pipeline_enable(True, lookahead=False, uncompacted=uncompacted)
            for k in ("susceptibility", "is_infected", "infection_time"):
                data["agent"][k] = state[k]
            data["agent"].symptoms = {k: state[k] for k in ("current_stage", "next_stage", "time_to_next_stage")}
            with ops.philox_seed(55):
                _, red = model.step(data, timer, age_bins=(0, 18, 65, 100))
            agent = data["agent"]
            outs[uncompacted] = dict(T=agent.transmission.detach().clone(), q=agent["not_infected_probs"].detach().clone(),
                                     n=agent["new_infected"].detach().clone(), red=red.detach().clone(),
                                     cur=agent.symptoms["current_stage"].detach().clone())
    finally:
        _lib.pipeline_enable(*prev)
    a, b = outs[True], outs[False]
    assert int((a["T"] != 0).sum()) > 1000
    assert torch.equal(a["T"], b["T"])
    rel = ((a["q"] - b["q"]).abs() / a["q"].clamp_min(1e-30)).max()
    assert float(rel) < 2e-6, float(rel)
    flips = int((a["n"] != b["n"]).sum())
    assert flips <= 2, flips
    assert torch.allclose(a["red"], b["red"], atol=flips + 0.5)
    H.report("compacted vs uncompacted transmission pass: draw flips", flips)


def test_throughput_mode_bptt_window():
    """Five timesteps of Runner() + backward (BPTT through the fused step) in throughput mode against the
    reference-order kernels on the same Philox stream: identical trajectories (unless a near-tie flips an
    agent) and matching log-beta gradients."""
    from grad_june import GradJune, Timer, ops
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    n_agents = 200_000
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = 5
    params["infection_seed"]["log_fraction_initial_cases"] = -1.5
    params["policies"] = {}
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=5, device=DEV))
    model = GradJune.from_parameters(params)
    keys = list(model.infection_networks.networks.keys())
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test", parameters=params)
    outs = {}
    for mode in ("exact", "fast"):
        leaves = []
        for k in keys:
            leaf = torch.tensor(float(params["networks"][k]["log_beta"]) + 0.4, device=DEV, requires_grad=True)
            model.infection_networks.networks[k].log_beta = leaf
            leaves.append(leaf)
        ops.EXACT_ORDER = mode == "exact"
        try:
            with ops.philox_seed(2024):
                results, is_inf = runner()
            loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum() \
                + 0.5 * results["cases_by_age_65"].sum()
            loss.backward()
        finally:
            ops.EXACT_ORDER = False
        outs[mode] = dict(cases=results["cases_per_timestep"].detach().cpu().numpy(), inf=is_inf.detach().cpu().numpy(),
                          cur=data["agent"].symptoms["current_stage"].detach().cpu().numpy(),
                          grads=torch.stack([l.grad for l in leaves]).cpu().numpy())
    e, f = outs["exact"], outs["fast"]
    assert e["cases"][-1] > e["cases"][0] > 0
    # unconditional: for this (world, seed) no draw of the window is a near-tie, so the two families must agree
    # exactly on the trajectory; if a kernel change makes this fail, teacher-forced tests tell a near-tie from a bug
    assert np.array_equal(e["inf"], f["inf"]), int((e["inf"] != f["inf"]).sum())
    assert np.array_equal(e["cases"], f["cases"])
    assert np.array_equal(e["cur"], f["cur"])
    assert np.allclose(e["grads"], f["grads"], rtol=5e-5, atol=1e-6 * np.abs(e["grads"]).max()), (e["grads"], f["grads"])


@pytest.mark.parametrize("quarantine", [False, True])
def test_pipelined_kernels_are_bit_identical(quarantine):
    """The TMA bulk-copy pipelined agent kernels (gj_pipe.cuh) against the register-batched ones (gj_lean.cuh):
    same arithmetic on the same Philox stream -> bit-identical trajectories and state; gradients to fp32 rounding.
    An odd agent count makes tiles start at unaligned agents (the copies start at the preceding 16-byte granule)."""
    from grad_june import GradJune, Timer, _lib, ops
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    n_agents = 333_337
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = 4
    params["infection_seed"]["log_fraction_initial_cases"] = -1.5
    params["policies"] = {}
    if quarantine:
        params["policies"] = {"quarantine": {"quarantine": {1: {"start_date": "2022-01-01", "end_date": "2023-01-01",
                                                                 "stage_threshold": 4}}}}
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=6, device=DEV, agents_per_super_area=5000))
    model = GradJune.from_parameters(params)
    keys = list(model.infection_networks.networks.keys())
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test", parameters=params)
    outs = {}
    prev = _lib.pipeline_enable(None)
    try:
        for mode in (False, True):
            _lib.pipeline_enable(mode, lookahead=False)
            leaves = []
            for k in keys:
                leaf = torch.tensor(float(params["networks"][k]["log_beta"]) + 0.4, device=DEV, requires_grad=True)
                model.infection_networks.networks[k].log_beta = leaf
                leaves.append(leaf)
            with ops.philox_seed(31):
                results, is_inf = runner()
            loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum() \
                + 0.5 * results["cases_by_age_65"].sum() + 0.01 * data["agent"].infection_time.sum()
            loss.backward()
            agent = data["agent"]
            outs[mode] = [results["cases_per_timestep"].detach(), results["deaths_per_timestep"].detach(), is_inf.detach(),
                          agent.susceptibility.detach(), agent.infection_time.detach(), agent.transmission.detach(),
                          agent.symptoms["current_stage"].detach(), agent.symptoms["next_stage"].detach(),
                          agent.symptoms["time_to_next_stage"].detach(), torch.stack([l.grad for l in leaves])]
    finally:
        _lib.pipeline_enable(*prev)
    assert outs[False][0][-1] > outs[False][0][0] > 0
    for a, b in zip(outs[False][:-1], outs[True][:-1]):
        assert torch.equal(a, b)
    # gradients: per-CTA partial sums are combined in grid order and the two families use different grids
    ga, gb = outs[False][-1], outs[True][-1]
    assert torch.allclose(ga, gb, rtol=2e-6, atol=1e-7 * float(ga.abs().max())), (ga, gb)


def test_graphed_runner_replays_the_eager_window():
    """GraphedRunner (Runner() + backward captured as one CUDA graph) against the eager Python loop on the same
    Philox key: identical results and gradients; a replay with other log-betas matches an eager run with them."""
    from grad_june import GradJune, Timer, ops
    from grad_june.default_config import default_parameters
    from grad_june.graphed import GraphedRunner
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    n_agents = 150_000
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = 5
    params["infection_seed"]["log_fraction_initial_cases"] = -1.5
    params["policies"] = {"quarantine": {"quarantine": {1: {"start_date": "2022-02-03", "end_date": "2023-01-01",
                                                             "stage_threshold": 4}}}}
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=8, device=DEV, agents_per_super_area=5000))
    model = GradJune.from_parameters(params)
    keys = list(model.infection_networks.networks.keys())
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test", parameters=params)
    loss_fn = lambda r: r["cases_per_timestep"].sum() + r["deaths_per_timestep"].sum() + 0.5 * r["cases_by_age_65"].sum()  # noqa: E731

    def eager(lb):
        leaves = []
        for i, k in enumerate(keys):
            leaf = lb[i].detach().clone().requires_grad_(True)
            model.infection_networks.networks[k].log_beta = leaf
            leaves.append(leaf)
        with ops.philox_seed(123):
            results, is_inf = runner()
        loss = loss_fn(results)
        loss.backward()
        return (loss.detach().clone(), torch.stack([l.grad for l in leaves]), results["cases_per_timestep"].detach().clone(),
                is_inf.detach().clone())

    base = torch.tensor([float(params["networks"][k]["log_beta"]) + 0.4 for k in keys], device=DEV)
    other = base + torch.linspace(-0.2, 0.2, len(keys), device=DEV)
    e1, e2 = eager(base), eager(other)
    graphed = GraphedRunner(runner, loss_fn, seed=123)
    for lb, ref in ((base, e1), (other, e2), (base, e1)):
        loss, grads, results = graphed(lb)
        torch.cuda.synchronize()
        assert torch.equal(results["cases_per_timestep"], ref[2])
        assert torch.equal(graphed.is_infected, ref[3])
        assert torch.equal(loss, ref[0])
        assert torch.equal(grads, ref[1])
    assert e1[2][-1] > e1[2][0] > 0 and not torch.equal(e1[1], e2[1])


@pytest.mark.parametrize("policies", [False, True])
def test_lookahead_transmission_pass(policies):
    """gj_step_forward_next: step t's agent kernel also runs step t+1's transmission pass (one pass over the agents
    less per step).  Same arithmetic per agent; only the tile sums of the leisure channels are associated
    differently, so trajectories agree up to near-tie flips and gradients to fp32 rounding.  With policies the
    look-ahead has to follow the schedule (quarantine mask, closed venues, weekday/weekend tables of the NEXT step)."""
    from grad_june import GradJune, Timer, _lib, ops
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    n_agents = 250_003
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = 8          # crosses a weekend (2022-02-05/06)
    params["infection_seed"]["log_fraction_initial_cases"] = -1.5
    params["policies"] = {}
    if policies:
        params["policies"] = {
            "quarantine": {"quarantine": {1: {"start_date": "2022-02-03", "end_date": "2022-02-07", "stage_threshold": 4}}},
            "close_venue": {"close_venue": {1: {"start_date": "2022-02-04", "end_date": "2022-02-08",
                                                "names": ["school", "pub", "gym"]}}},
            "interaction": {"social_distancing": {1: {"start_date": "2022-02-02", "end_date": "2022-02-06",
                                                      "beta_factors": {"company": 0.5, "visit": 0.5}}}}}
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=7, device=DEV, agents_per_super_area=5000))
    model = GradJune.from_parameters(params)
    keys = list(model.infection_networks.networks.keys())
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test", parameters=params)
    outs = {}
    prev = _lib.pipeline_enable(None)
    try:
        for look in (False, True):
            _lib.pipeline_enable(True, lookahead=look)
            leaves = []
            for k in keys:
                leaf = torch.tensor(float(params["networks"][k]["log_beta"]) + 0.4, device=DEV, requires_grad=True)
                model.infection_networks.networks[k].log_beta = leaf
                leaves.append(leaf)
            _lib.profile_enable(True)
            with ops.philox_seed(32):
                results, is_inf = runner()
            loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum() \
                + 0.5 * results["cases_by_age_65"].sum()
            loss.backward()
            launches = _lib.profile_read()["transmission"][2]
            _lib.profile_enable(False)
            outs[look] = dict(cases=results["cases_per_timestep"].detach().cpu().numpy(), inf=is_inf.detach().cpu().numpy(),
                              T=data["agent"].transmission.detach().cpu().numpy(),
                              grads=torch.stack([l.grad for l in leaves]).cpu().numpy(), launches=launches)
    finally:
        _lib.pipeline_enable(*prev)
    a, b = outs[False], outs[True]
    assert a["launches"] == 8 and b["launches"] == 1          # only the first step still runs the stand-alone pass
    assert a["cases"][-1] > a["cases"][0] > 0
    assert np.array_equal(a["inf"], b["inf"]), int((a["inf"] != b["inf"]).sum())     # unconditional (see above)
    assert np.array_equal(a["cases"], b["cases"])
    assert np.array_equal(a["T"], b["T"])
    assert np.allclose(a["grads"], b["grads"], rtol=5e-5, atol=1e-6 * np.abs(a["grads"]).max()), (a["grads"], b["grads"])


def test_ensemble_evaluator_matches_single_runs():
    """EnsembleEvaluator (calibration ensemble, one graph replay per sample) against eager runs of the same samples."""
    from grad_june import GradJune, Timer, ops
    from grad_june.calibration import EnsembleEvaluator
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    n_agents = 60_000
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = 3
    params["infection_seed"]["log_fraction_initial_cases"] = -1.5
    params["policies"] = {}
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=9, device=DEV, agents_per_super_area=5000))
    model = GradJune.from_parameters(params)
    keys = list(model.infection_networks.networks.keys())
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test", parameters=params)
    loss_fn = lambda r: r["cases_per_timestep"].sum() + r["deaths_per_timestep"].sum()  # noqa: E731
    base = torch.tensor([float(params["networks"][k]["log_beta"]) + 0.4 for k in keys], device=DEV)
    samples = base + 0.25 * torch.randn(3, len(keys), generator=torch.Generator().manual_seed(1)).to(DEV)
    ref_loss, ref_grads = [], []
    for lb in samples:
        leaves = []
        for i, k in enumerate(keys):
            leaf = lb[i].detach().clone().requires_grad_(True)
            model.infection_networks.networks[k].log_beta = leaf
            leaves.append(leaf)
        with ops.philox_seed(11):
            results, _ = runner()
        loss = loss_fn(results)
        loss.backward()
        ref_loss.append(loss.detach())
        ref_grads.append(torch.stack([l.grad for l in leaves]))
    losses, grads = EnsembleEvaluator(runner, loss_fn, seed=11)(samples)
    assert torch.equal(losses, torch.stack(ref_loss))
    assert torch.equal(grads, torch.stack(ref_grads))
    assert not torch.equal(grads[0], grads[1])


def test_calibrator_recovers_a_shifted_log_beta(tmp_path):
    """Calibrator (optimiser loop over the captured window): fit one network's log-beta to the time series
    the model itself produced at another value; the loss must fall by orders of magnitude and the parameter move
    towards the truth.  Also the CSV outputs of save_results (runner.py:185-196) and the one-pass ethnicity reduction."""
    import pandas as pd
    from grad_june import GradJune, Timer, ops
    from grad_june.calibration import Calibrator
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    n_agents = 50_000
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = 8
    params["infection_seed"]["log_fraction_initial_cases"] = -2.0
    params["policies"] = {}
    params["save_path"] = str(tmp_path / "out")
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=3, device=DEV, agents_per_super_area=5000))
    data["agent"].ethnicity = np.array(["A", "B", "C"])[np.arange(n_agents) % 3]
    model = GradJune.from_parameters(params)
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-2.0,
                    save_path=params["save_path"], parameters=params)
    nets = model.infection_networks.networks
    truth = float(nets["household"].log_beta) + 0.5
    start = float(nets["household"].log_beta)
    nets["household"].log_beta = torch.tensor(truth)
    with torch.no_grad(), ops.philox_seed(3):
        target = runner()[0]["cases_per_timestep"].detach().clone()
    by_eth = runner.get_cases_by_ethnicity(runner.data)
    inf = runner.data["agent"].is_infected
    codes = torch.arange(n_agents, device=DEV) % 3
    assert torch.equal(by_eth, torch.stack([(inf * (codes == k)).sum() for k in range(3)]))
    nets["household"].log_beta = torch.tensor(start)
    scale = float(target.max())
    cal = Calibrator(runner, loss_fn=lambda r: (((r["cases_per_timestep"] - target) / scale) ** 2).mean(),
                     networks=["household"], lr=0.1, seed=3)
    history = cal.fit(30)
    assert history["loss"].iloc[-1] < 0.05 * history["loss"].iloc[0]
    assert abs(history["log_beta_household"].iloc[-1] - truth) < 0.5 * abs(start - truth)
    path = cal.save(history)
    res = pd.read_csv(path / "results.csv")
    assert list(res.columns[:4]) == ["date", "cases_per_timestep", "daily_cases_per_timestep", "deaths_per_timestep"]
    assert len(res) == 9 and len(pd.read_csv(path / "results_is_infected.csv")) == n_agents
    assert len(pd.read_csv(path / "calibration.csv")) == 30
