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


# ------------------------------------------------------------------------------------------
# gradient parity criterion
# ------------------------------------------------------------------------------------------
class perturb_q:
    """Context manager: every q the oracle computes is moved by a random -1/0/+1 fp32 ulp (values only,
    same autograd graph).  Measures how much of a parameter gradient is decided by the last bit of
    expf — the part that legitimately differs between CPU libm (reference) and CUDA libm (kernels)."""

    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        gen = torch.Generator().manual_seed(self.seed)
        self.orig = orig = O.not_infected_probs

        def hooked(world, spec, T, s, cur, return_pressure=False):
            q, lam = orig(world, spec, T, s, cur, True)
            if q.dtype == torch.float32:
                r = (torch.randint(0, 3, q.shape, generator=gen) - 1).to(q.device)
                qp = (q.detach().view(torch.int32) + r.int()).view(torch.float32).clamp(max=1.0)
                q = q + (qp - q.detach())
            return (q, lam) if return_pressure else q
        O.not_infected_probs = hooked

    def __exit__(self, *a):
        O.not_infected_probs = self.orig


def oracle_run(tag, dtype=torch.float32, device="cpu", noises=None, loss_weights=None):
    """Replay a golden Runner trajectory through the oracle; returns (result dict, grads[11], grad log_frac,
    masks_equal_to_golden).  ``noises``: a list of StepNoise replacing the golden run's injected noise (e.g. the
    kernels' own Philox stream from gj_philox_fill); the last return value then compares with nothing useful."""
    from grad_june.policies import Policies
    from grad_june.symptoms import SymptomsSampler
    g = np.load(GOLDEN / f"run_{tag}.npz")
    params, _ = load_params(tag)
    arrays = np.load(GOLDEN / "sample_world.npz")
    w = oracle_world(arrays, SAMPLE_TYPES, device)
    nets = make_leaf_networks(params)
    policies = Policies.from_parameters(params)
    sym = oracle_symptoms(SymptomsSampler.from_parameters(params), device)
    steps = oracle_schedule(params, nets, policies)
    if dtype != torch.float32:
        for s in steps:
            for n in s.nets:
                n.beta = n.beta.to(dtype)
    noises = [O.StepNoise(E=n.E.to(device=device, dtype=dtype), u=n.u.to(device=device, dtype=dtype),
                          z=n.z.to(device=device, dtype=dtype))
              for n in (torch_noise(RUNS[tag], len(steps) + 1, w.n_agents, device) if noises is None else noises)]
    log_frac = torch.tensor(float(params["infection_seed"]["log_fraction_initial_cases"]), requires_grad=True,
                            dtype=dtype)
    prof = {k: v.to(dtype) for k, v in profile_params(g, device).items()}
    trace = []
    res = O.run(w, prof, sym, log_frac, steps, noises, age_bins=params.get("age_bins_to_save", (0, 18, 65, 100)),
                dtype=dtype, trace=trace)
    wc, wd, wa = g["loss_weights"] if loss_weights is None else loss_weights
    cba = res["cases_by_age"]
    res["trace"] = trace
    loss = wc * res["cases_per_timestep"].sum() + wd * res["deaths_per_timestep"].sum() \
        + wa * (cba * torch.arange(1, cba.shape[1] + 1, device=device)).sum()
    loss.backward()
    names = [str(n) for n in g["net_names"]]
    grads = np.array([nets.networks[n].log_beta.grad.item() if nets.networks[n].log_beta.grad is not None else 0.0
                      for n in names])
    same = np.array_equal(np.stack([t["is_infected"].cpu().numpy() for t in trace]).astype(np.uint8),
                          g["trace_is_infected"]) and \
        np.array_equal(np.stack([t["current_stage"].cpu().numpy() for t in trace]).astype(np.uint8),
                       g["trace_current_stage"])
    return res, grads, log_frac.grad.item(), same


def assert_grad_parity(mine, ref32, f64, sens=None, rtol=1e-5, slack=4.0, ulp_slack=6.0, what="gradient"):
    """Parameter-gradient criterion.

    Target: 1e-5 relative agreement with the reference's fp32 gradient.  That is only meaningful where the
    reference's gradient is itself conditioned to 1e-5: the factor y0*y1*dt/(tau*(1-q)) turns a 1-ulp
    difference in q (CPU expf vs CUDA expf) into a relative change ulp/(1-q) of the term, i.e. 1e-4..1e-2
    for agents under a small pressure.  A component therefore passes if
      (a) it is within ``rtol`` of the reference, or
      (b) its distance to the fp64 witness of the same trajectory is at most ``slack`` x the reference's own
          distance to that witness (not less accurate than the reference's fp32 arithmetic), or
      (c) its distance to the reference is at most ``ulp_slack`` x the change the reference's own gradient
          shows when its q values are moved by a random +-1 ulp (``sens``, measured with perturb_q)."""
    mine, ref32, f64 = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (mine, ref32, f64))
    sens = np.zeros_like(ref32) if sens is None else np.atleast_1d(np.asarray(sens, dtype=np.float64))
    used = {"a": 0, "b": 0, "c": 0, "max_rel": 0.0}
    for i, (m, r, t, sn) in enumerate(zip(mine, ref32, f64, sens)):
        if abs(r) > 0:
            used["max_rel"] = max(used["max_rel"], abs(m - r) / abs(r))
        if abs(m - r) <= rtol * abs(r) + 1e-30:
            used["a"] += 1
            continue
        e_ref, e_mine = abs(r - t), abs(m - t)
        if e_mine <= max(slack * e_ref, rtol * abs(t)):
            used["b"] += 1
            continue
        assert abs(m - r) <= ulp_slack * sn, \
            (f"{what}[{i}]: mine {m!r} ref32 {r!r} fp64 {t!r}: |mine-ref| = {abs(m - r):.3e} (rel {abs(m - r) / abs(r):.2e}), "
             f"|mine-f64| = {e_mine:.3e}, |ref32-f64| = {e_ref:.3e}, 1-ulp sensitivity = {sn:.3e}")
        used["c"] += 1
    report(what, used)
    return used


_REPORT = {}


def report(key, value):
    """Collected per test session and written to gpurun_out/parity_report.json (see conftest.py): which gradient
    components needed the fall-back criteria, how many masks were near-ties, the largest gaps."""
    _REPORT.setdefault(str(key), []).append(value)


def run_sensitivity(tag, seeds=(1, 2, 3), **kw):
    """max over seeds of |grad(perturbed q) - grad| for the golden run ``tag`` (fp32 oracle, CPU)."""
    _, base, basef, _ = oracle_run(tag, **kw)
    sens, sensf = np.zeros_like(base), 0.0
    for sd in seeds:
        with perturb_q(sd):
            _, gp, gpf, _ = oracle_run(tag, **kw)
        sens = np.maximum(sens, np.abs(gp - base))
        sensf = max(sensf, abs(gpf - basef))
    return sens, sensf


def oracle_step100(dtype=torch.float32, device="cpu"):
    """The single-step fixture through the oracle: returns (aux dict, post state, grads[household, company, school])."""
    from grad_june.symptoms import SymptomsSampler
    g = np.load(GOLDEN / "step100.npz")
    w = oracle_world(g, ["school", "company", "household"], device)
    state = {k: torch.from_numpy(g["pre_" + k]).to(device=device, dtype=dtype) for k in
             ("susceptibility", "is_infected", "infection_time", "current_stage", "next_stage", "time_to_next_stage")}
    params = {k: v.to(dtype) for k, v in profile_params(g, device).items()}
    sym = oracle_symptoms(SymptomsSampler.from_file(), device)
    lb = {"household": torch.tensor(0.5, requires_grad=True, dtype=dtype),
          "company": torch.tensor(0.3, requires_grad=True, dtype=dtype),
          "school": torch.tensor(0.4, requires_grad=True, dtype=dtype)}
    kinds = {"household": O.KIND_HOUSEHOLD, "company": O.KIND_PLAIN, "school": O.KIND_PLAIN}
    nets = [O.NetSpec(str(n), str(n), kinds[str(n)], 10.0 ** lb[str(n)]) for n in g["order"]]
    spec = O.StepSpec(now=float(g["now"]), dt=float(g["dt"]), day_type=0, nets=nets, quarantine=[])
    nz = torch_noise(8, 1, w.n_agents, device)[0]
    nz = O.StepNoise(E=nz.E.to(dtype), u=nz.u.to(dtype), z=nz.z.to(dtype))
    aux = {}
    O.step(w, state, params, spec, sym, nz, aux)
    wl = torch.from_numpy(g["loss_w"]).to(device=device, dtype=dtype)
    w2 = torch.from_numpy(g["loss_w2"]).to(device=device, dtype=dtype)
    loss = (state["is_infected"] * wl).sum() + (state["current_stage"] * w2).sum() \
        + 0.5 * (state["susceptibility"] * w2).sum() + 0.1 * (state["infection_time"] * wl).sum()
    loss.backward()
    return aux, state, np.array([lb[k].grad.item() for k in ("household", "company", "school")])


# ------------------------------------------------------------------------------------------
# Philox-mode parity: the kernels' own noise for the oracle, and the near-tie certificate
# ------------------------------------------------------------------------------------------
def philox_noises(seed, n_calls, n_agents, device="cpu", first_call=0, gpu="cuda:0"):
    """The kernels' own draws for calls first_call .. first_call + n_calls - 1 and the LOADED agent ids 0..n-1, as
    the oracle's injected noise (one normal per agent and call: every dwell-time row reads the same draw)."""
    from grad_june import ops
    out = []
    for c in range(first_call, first_call + n_calls):
        E, u, z = ops.philox_fill(seed, c, n_agents, gpu)
        out.append(O.StepNoise(E=E.to(device), u=u.to(device), z=z.to(device).expand(10, n_agents)))
    return out


def certify_near_ties(q, E, mism, what=""):
    """Every mask mismatch must be a draw the last bits decide.  q: the oracle's fp32 not-infected probabilities,
    E[2, n]: the exponentials, mism: indices.  The draw is sign(d), d = (ln q - ln E0) - (ln(1-q) - ln E1).  The
    throughput kernels evaluate d with six hardware log2 (lg2.approx: absolute error <= 2^-22 for arguments in
    (0.5, 2), relative 2^-22 elsewhere), the reference with six correctly rounded logf plus two roundings: both carry
    an absolute error of a few 2^-22 x the largest term.  A mismatch is certified when |d| (fp64, from the fp32
    inputs) is within 8 x 2^-22 x max(1, |terms|), or when moving q by one fp32 ulp — the accuracy of expf, which
    (1 - q) amplifies by 1 / (1 - q) for agents under tiny pressure — changes the sign of d."""
    if len(mism) == 0:
        return 0.0
    q32 = np.asarray(q, dtype=np.float32)[mism]
    E0, E1 = (np.asarray(E[i], dtype=np.float64)[mism] for i in (0, 1))

    def d_of(qq):
        qq = qq.astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (np.log(qq) - np.log(E0)) - (np.log1p(-qq) - np.log(E1))

    with np.errstate(divide="ignore", invalid="ignore"):
        terms = np.maximum.reduce([np.ones_like(E0), np.abs(np.log(q32.astype(np.float64))),
                                   np.abs(np.log1p(-q32.astype(np.float64))), np.abs(np.log(E0)), np.abs(np.log(E1))])
    d = d_of(q32)
    lo = d_of(np.nextafter(q32, np.float32(0.0)))
    hi = d_of(np.minimum(np.nextafter(q32, np.float32(2.0)), np.float32(1.0)))
    tol = 8.0 * 2.0 ** -22 * terms
    ok = (np.abs(d) <= tol) | (np.sign(lo) != np.sign(hi)) | (np.sign(lo) != np.sign(d))
    assert ok.all(), f"{what}: mask mismatch that is not a near-tie: d = {d[~ok]}, tol = {tol[~ok]}, q = {q32[~ok]}"
    return float(np.max(np.abs(d) / tol))


def mid_epidemic_state(n, now, seed, device):
    """A plausible state in the middle of an epidemic (a quarter of the agents infected at various stages, some
    recovered / dead), generated ON ``device`` so that it also serves the 56 M-agent verification of bench.py."""
    g = torch.Generator(device=device).manual_seed(seed)

    def rand():
        return torch.rand(n, generator=g, device=device)

    inf = rand() < 0.25
    zeros, ones = torch.zeros(n, device=device), torch.ones(n, device=device)
    stage = torch.randint(2, 7, (n,), generator=g, device=device).float()
    cur = torch.where(inf, stage, ones)
    nxt = torch.where(inf, torch.where(rand() < 0.4, zeros, stage + 1), ones)
    done = inf & (rand() < 0.2)                        # already recovered / dead
    cur = torch.where(done, torch.where(rand() < 0.9, zeros, torch.full_like(zeros, 7.0)), cur)
    nxt = torch.where(done, cur, nxt)
    tinf = torch.where(inf, now - 12.0 * rand(), zeros)
    ttn = torch.where(inf, now + 4.0 * rand() - 1.5, zeros)
    s = torch.where(inf, zeros, ones)
    return {"susceptibility": s, "is_infected": inf.float(), "infection_time": tinf, "current_stage": cur,
            "next_stage": nxt, "time_to_next_stage": ttn}


def oracle_step_inputs(params, data, model, timer, state, device):
    """Oracle inputs of ONE step of ``model`` on ``data`` (in data's own agent numbering) at ``timer``:
    (world, leaf networks, StepSpec, SymptomsSpec, profile, state copy)."""
    w = O.OracleWorld(n_agents=len(data["agent"].id), age=data["agent"].age.to(device), sex=data["agent"].sex.to(device))
    for t in data.venue_types():
        ei = data["attends_" + t].edge_index.to(device)
        w.edges[t] = O.EdgeType(src=ei[0], dst=ei[1], people=torch.as_tensor(data[t]["people"]).to(device),
                                n_groups=len(data[t]["id"]))
    nets = make_leaf_networks({**params, "system": {"device": "cpu"}})
    with torch.no_grad():
        for k, net in nets.networks.items():
            net.log_beta.copy_(torch.as_tensor(model.infection_networks.networks[k].log_beta).detach().cpu())
    policies = model.policies
    specs = []
    for net in nets.active_networks(timer, policies):
        prob = getattr(net, "leisure_probabilities", None)
        specs.append(O.NetSpec(net.name, net.edge_type(), net.kind, net.beta_eff(policies, timer).to(device),
                               None if prob is None else prob.to(device)))
    quar = policies.quarantine_policies.active_thresholds(timer) if policies.quarantine_policies else None
    spec = O.StepSpec(now=timer.now, dt=timer.duration, day_type=0 if timer.day_type == "weekday" else 1, nets=specs,
                      quarantine=quar)
    sym = oracle_symptoms(SymptomsSampler.from_parameters(params), device)
    prof = {k: v.to(device) for k, v in data["agent"].infection_parameters.items()}
    st = {k: v.detach().clone().to(device) for k, v in state.items()}
    return w, nets, spec, sym, prof, st


def layout_noise(data, E, u, z, device):
    """The kernels' Philox draws for LOADED ids 0..n-1 (ops.philox_fill) as the oracle's StepNoise in ``data``'s
    own numbering (a renumbered world reads agent i's draw at original_index[i])."""
    agent = data["agent"]
    if "original_index" in agent:
        oi = agent["original_index"].to(E.device)
        E, u, z = E[:, oi], u[oi], z[oi]
    n = u.numel()
    return O.StepNoise(E=E.to(device), u=u.to(device), z=z.to(device).expand(10, n))
