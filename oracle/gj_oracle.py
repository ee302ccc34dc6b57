"""ORACLE — test infrastructure, NOT product code.

A flat, functional torch restatement of GradABM-JUNE's per-timestep infection path with every
random draw taken from explicit ("injected") noise tensors.  It is the checker the CUDA path is
compared against: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package never does.

Why torch: the path is floating point and differentiable; the reference *is* a sequence of torch
fp32 ops whose gradients come from autograd.  Restating it op-for-op in torch (same op order, same
dtype promotions) gives bit-comparable forward values on CPU and autograd gradients for free.  The
same code runs in fp64 (``dtype=torch.float64``) as a higher-precision witness and on ``cuda`` as a
same-device witness.

Parity pinning: ``tests/golden/make_golden.py`` runs the reference's own code (imported from
/root/reference behind ``tests/golden/_pyg_shim.py``) with the same injected noise and commits
its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
them bit-for-bit (forward) / 1e-6 (gradients).  The reference's KAT
(test/unit/infection_networks/test_base.py:39-44) is reproduced in the same test file.

Each function cites the reference file:line it restates (paths relative to /root/reference).
Third-party piece restated: torch_geometric ``MessagePassing.propagate`` with aggr="add"
(requirements.txt:5, ">=2.3", not vendored) == gather, multiply, ``index_add_`` in edge order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

TAU = 0.1  # infection.py:15 (gumbel_softmax tau)

# network "kinds": how the per-agent masks are formed (base.py:47-59,144-149;
# leisure_network.py:61-85,107-120)
KIND_PLAIN = 0       # mT = mS = quarantine mask
KIND_HOUSEHOLD = 1   # mT = mS = 1  (household ignores quarantine)
KIND_LEISURE = 2     # mT = mS = quarantine mask * prob[daytype, sex, age]
KIND_CARE_VISIT = 3  # as leisure, susceptibility additionally * (age > 75)


@dataclass
class EdgeType:
    """One agent->group edge set, reference layout (june_world_loader/graph_loader.py:16-39)."""
    src: torch.Tensor      # [E] int64 agent ids, reference (unsorted) order
    dst: torch.Tensor      # [E] int64 group ids
    people: torch.Tensor   # [G] int64 or float: data[name]["people"]
    n_groups: int


@dataclass
class OracleWorld:
    n_agents: int
    age: torch.Tensor                   # [N] int64 in [0, 99]
    sex: torch.Tensor                   # [N] int64 in {0, 1}
    edges: Dict[str, EdgeType] = field(default_factory=dict)


@dataclass
class NetSpec:
    """One active infection network for one step (already ordered by timer.py:14-26)."""
    name: str
    edge_type: str                      # "leisure" for all LeisureNetwork subclasses (leisure_network.py:44-48)
    kind: int
    beta: torch.Tensor                  # scalar beta_eff = 10**log_beta * prod(policy factors) (base.py:36-42)
    prob: Optional[torch.Tensor] = None  # [2, 2, 100] leisure table (leisure_network.py:26-34)


@dataclass
class StepSpec:
    now: float                          # timer.now, days
    dt: float                           # timer.duration, days
    day_type: int                       # 0 weekday / 1 weekend (timer.py:84-90)
    nets: List[NetSpec]
    quarantine: Optional[List[float]]   # None: no quarantine collection; else thresholds of ACTIVE policies


@dataclass
class SymptomsSpec:
    """symptoms.py:10-63."""
    n_stages: int
    prob: torch.Tensor                  # [S, 100] stage transition probabilities by age
    # per stage i: (dist_kind, loc, scale) or None; dist_kind 0 = LogNormal, 1 = Normal
    trans_times: Dict[int, Optional[tuple]] = field(default_factory=dict)
    rec_times: Dict[int, Optional[tuple]] = field(default_factory=dict)


@dataclass
class StepNoise:
    E: torch.Tensor                     # [2, N] Exp(1) draws (torch functional.py:2218-2222)
    u: torch.Tensor                     # [N] U[0,1) for the Bernoulli branch (symptoms.py:97)
    z: torch.Tensor                     # [2*(S-3), N] standard normals, row (i-2)*2 + {0: progress, 1: recover}


def _f(x, like: torch.Tensor):
    return torch.as_tensor(x, dtype=like.dtype, device=like.device)


# --------------------------------------------------------------------------------------
# a1  transmission.py:38-51
# --------------------------------------------------------------------------------------
def transmission(now: float, infection_time, is_infected, max_infectiousness, shape, rate, shift):
    t = now - infection_time
    sign = (torch.sign(t - shift + 1e-10) + 1) / 2
    aux = torch.exp(-torch.lgamma(shape)) * torch.pow((t - shift) * rate, shape - 1.0)
    aux2 = torch.exp((shift - t) * rate) * rate
    return max_infectiousness * sign * aux * aux2 * is_infected


# --------------------------------------------------------------------------------------
# a2  policies/quarantine_policies.py:13-33
# --------------------------------------------------------------------------------------
def quarantine_mask(current_stage, thresholds: Optional[Sequence[float]]):
    """None -> scalar 1.0 (base.py:48-51 'else' branch); [] -> ones; else product of (stage < thr)."""
    if thresholds is None:
        return 1.0
    m = torch.ones(current_stage.shape, device=current_stage.device)
    for thr in thresholds:
        m = m * (current_stage < thr).to(torch.float)
    return m.to(current_stage.dtype) if current_stage.dtype == torch.float64 else m


# --------------------------------------------------------------------------------------
# a6  torch_geometric propagate(aggr="add"), called at base.py:79-83 with message base.py:86-87
# --------------------------------------------------------------------------------------
def propagate(src, dst, x, y):
    msg = x.index_select(-1, src) * y.index_select(-1, dst)
    return torch.zeros(y.shape[-1], dtype=msg.dtype, device=msg.device).index_add(0, dst, msg)


def p_contact(people, like):
    """base.py:64-69 ; people may be int64 (pickles) or float (test fixtures)."""
    return torch.maximum(torch.minimum(1.0 / (people - 1), _f(1.0, like)), _f(0.0, like)).to(like.dtype)


# --------------------------------------------------------------------------------------
# a5/a7  base.py:61-84 + mask overrides
# --------------------------------------------------------------------------------------
def network_pressure(world: OracleWorld, net: NetSpec, T, s, qmask, day_type: int):
    et = world.edges[net.edge_type]
    beta = net.beta * torch.ones(et.n_groups, dtype=T.dtype, device=T.device)   # base.py:41
    beta = beta * p_contact(et.people.to(T.device), T)                          # base.py:70
    if net.kind == KIND_HOUSEHOLD:                                              # base.py:144-149
        Tm, sm = T, s
    elif net.kind == KIND_PLAIN:                                                # base.py:47-59
        Tm, sm = qmask * T, qmask * s
    else:                                                                       # leisure_network.py:61-85
        lm = net.prob.to(T.dtype)[day_type, world.sex, world.age]
        Tm = qmask * lm * T
        sm = qmask * lm * s
        if net.kind == KIND_CARE_VISIT:                                         # leisure_network.py:108-120
            sm = sm * (world.age > 75)
    cum = propagate(et.src, et.dst, Tm, beta)                                   # base.py:79
    return propagate(et.dst, et.src, cum, sm)                                   # base.py:80-83


# --------------------------------------------------------------------------------------
# a3  base.py:118-141
# --------------------------------------------------------------------------------------
def not_infected_probs(world, spec: StepSpec, T, s, current_stage, return_pressure=False):
    qmask = quarantine_mask(current_stage, spec.quarantine)
    lam = torch.zeros(world.n_agents, dtype=T.dtype, device=T.device)
    for net in spec.nets:
        lam = lam + network_pressure(world, net, T, s, qmask, spec.day_type)
    lam_raw = lam
    lam = torch.clamp(lam, min=1e-6, max=100)
    q = torch.exp(-lam * spec.dt)
    q = torch.clamp(q, min=0.0, max=1.0)
    if return_pressure:
        return q, lam_raw
    return q


# --------------------------------------------------------------------------------------
# a8  infection.py:3-18 + torch.nn.functional.gumbel_softmax (functional.py:2165-2235), noise injected
# --------------------------------------------------------------------------------------
def is_infected_sample(q, E, return_soft=False):
    logits = torch.vstack((q, 1.0 - q)).log()
    gumbels = -E.log()
    gumbels = (logits + gumbels) / TAU
    y_soft = gumbels.softmax(0)
    index = y_soft.max(0, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(0, index, 1.0)
    ret = y_hard - y_soft.detach() + y_soft
    n = 1.0 - ret[0, :]
    if return_soft:
        return n, y_soft
    return n


# --------------------------------------------------------------------------------------
# a9  model.py:90-110 (maximum variant) / infection.py:21-28 (clamp variant, seeding)
# --------------------------------------------------------------------------------------
def infect_people(state: dict, n, now: float, clamp_variant=False):
    s = state["susceptibility"]
    if clamp_variant:
        state["susceptibility"] = torch.clamp(s - n, min=0.0)
    else:
        state["susceptibility"] = torch.maximum(_f(0.0, s), s - n)
    state["is_infected"] = state["is_infected"] + n
    state["infection_time"] = state["infection_time"] + n * (now - state["infection_time"])


def _dwell(spec_entry, z, like):
    kind, loc, scale = spec_entry
    x = _f(loc, like) + z * _f(scale, like)      # Normal.rsample: loc + eps * scale
    return torch.exp(x) if kind == 0 else x      # LogNormal = ExpTransform(Normal)


# --------------------------------------------------------------------------------------
# a10  symptoms.py:204-247 + 82-128  (the `if n_symp > 0` gates only skip adding zeros)
# --------------------------------------------------------------------------------------
def symptoms_update(state: dict, n, now: float, age, sym: SymptomsSpec, u, z):
    like = state["time_to_next_stage"]
    nxt = state["next_stage"] + n * (2.0 - state["next_stage"])
    ttn = state["time_to_next_stage"] + n * (now - state["time_to_next_stage"])
    cur = state["current_stage"]
    mask_transition = (now >= ttn) * (cur < sym.n_stages - 1)
    cur = cur - (cur - nxt) * mask_transition
    probs = sym.prob.to(like.dtype)[cur.long(), age]
    mask_symp_stage = u < probs
    mask_rec_stage = ~mask_symp_stage
    for i in range(2, sym.n_stages - 1):
        mask_stage = cur == i
        mask_stage = mask_stage * cur / i
        mask_updating = mask_stage * mask_transition
        mask_symp = mask_updating * mask_symp_stage
        if sym.trans_times.get(i) is not None:
            nxt = nxt + mask_symp
            ttn = ttn + _dwell(sym.trans_times[i], z[(i - 2) * 2], like) * mask_symp
        mask_rec = mask_updating * mask_rec_stage
        if sym.rec_times.get(i) is not None:
            nxt = nxt - nxt * mask_rec
            ttn = ttn + _dwell(sym.rec_times[i], z[(i - 2) * 2 + 1], like) * mask_rec
    state["current_stage"] = cur
    state["next_stage"] = nxt
    state["time_to_next_stage"] = ttn


# --------------------------------------------------------------------------------------
# a11  model.py:112-144
# --------------------------------------------------------------------------------------
def step(world, state: dict, params: dict, spec: StepSpec, sym: SymptomsSpec, noise: StepNoise, aux=None):
    """state: susceptibility,is_infected,infection_time,current_stage,next_stage,time_to_next_stage.
    params: max_infectiousness, shape, rate, shift."""
    T = transmission(spec.now, state["infection_time"], state["is_infected"],
                     params["max_infectiousness"], params["shape"], params["rate"], params["shift"])
    state["transmission"] = T
    q, lam = not_infected_probs(world, spec, T, state["susceptibility"], state["current_stage"], True)
    n, y = is_infected_sample(q, noise.E, True)
    if aux is not None:
        aux.update(q=q, lam=lam, new_infected=n, y_soft=y, transmission=T)
    infect_people(state, n, spec.now, clamp_variant=False)
    symptoms_update(state, n, spec.now, world.age, sym, noise.u, noise.z)
    return state


# --------------------------------------------------------------------------------------
# a12  runner.py:138-149 + infection.py:31-42  (seeding) and runner.py:198-224 (reductions)
# --------------------------------------------------------------------------------------
def seed(world, state, log_fraction, now, sym, noise: StepNoise, aux=None):
    like = state["susceptibility"]
    fraction = 10.0 ** log_fraction
    probs = fraction * torch.ones(world.n_agents, dtype=like.dtype, device=like.device)
    n, y = is_infected_sample(1.0 - probs, noise.E, True)
    if aux is not None:
        aux.update(new_infected=n, y_soft=y)
    infect_people(state, n, now, clamp_variant=True)
    symptoms_update(state, n, now, world.age, sym, noise.u, noise.z)
    return state


def deaths(state, sym: SymptomsSpec):
    dead = sym.n_stages - 1
    cur = state["current_stage"]
    return ((cur == dead) * cur / dead).sum()


def cases_by_age(world, state, age_bins):
    out = []
    for i in range(1, len(age_bins)):
        mask = (world.age < age_bins[i]) * (world.age > age_bins[i - 1])
        out.append((state["is_infected"] * mask).sum())
    return torch.stack(out)


def initial_state(n, dtype=torch.float32, device="cpu"):
    """runner.py:81-90 (stages start as int64 ones, become float after the first update)."""
    return {
        "susceptibility": torch.ones(n, dtype=dtype, device=device),
        "is_infected": torch.zeros(n, dtype=dtype, device=device),
        "infection_time": torch.zeros(n, dtype=dtype, device=device),
        "current_stage": torch.ones(n, dtype=torch.long, device=device),
        "next_stage": torch.ones(n, dtype=torch.long, device=device),
        "time_to_next_stage": torch.zeros(n, dtype=dtype, device=device),
    }


def run(world, params, sym, log_fraction, steps: List[StepSpec], noises: List[StepNoise],
        age_bins=(0, 18, 65, 100), dtype=torch.float32, state=None, trace=None):
    """runner.py:151-183.  noises[0] is the seeding draw, noises[1+t] belongs to steps[t].
    The seeding happens at now = 0 (timer.reset, runner.py:155)."""
    st = initial_state(world.n_agents, dtype, world.age.device) if state is None else state
    seed(world, st, log_fraction, 0.0, sym, noises[0])
    cases = [st["is_infected"].sum()]
    dts = [deaths(st, sym)]
    cba = [cases_by_age(world, st, age_bins)]
    if trace is not None:
        trace.append({k: v.detach().clone() for k, v in st.items()})
    for spec, noise in zip(steps, noises[1:]):
        step(world, st, params, spec, sym, noise)
        cases.append(st["is_infected"].sum())
        dts.append(deaths(st, sym))
        cba.append(cases_by_age(world, st, age_bins))
        if trace is not None:
            trace.append({k: v.detach().clone() for k, v in st.items()})
    return {
        "cases_per_timestep": torch.stack(cases),
        "deaths_per_timestep": torch.stack(dts),
        "cases_by_age": torch.stack(cba),
        "state": st,
    }
