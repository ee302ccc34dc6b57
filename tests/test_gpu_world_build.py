"""The C-ABI world builder on the GPU (gj_world_build) and a C-ABI-ONLY step: the reference's sample world is built,
renumbered and stepped through ctypes alone — no grad_june.world / grad_june.ops — the way a maintainer of the
reference would bind libgradjune_b200.so, and checked against the oracle on the world as loaded."""
import ctypes as C

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


def test_device_build_matches_python_builder():
    from grad_june import world as W
    from test_world_build import _compare
    n = 300_000
    data = W.make_synthetic_world(n, seed=12, device=DEV, agents_per_super_area=5000)
    data = W.renumber_world(data, torch.randperm(n, device=DEV))
    del data["agent"]["original_index"]
    native = W.NativeWorld(data, device=DEV, renumber=True)
    perm = native.permutation()
    data = W.renumber_world(data)
    assert torch.equal(perm, data["agent"].original_index)
    types = data.venue_types()
    ref = W.build_csr(n, types, {t: data["attends_" + t].edge_index for t in types},
                      {t: torch.as_tensor(data[t]["people"]) for t in types}, {t: len(data[t]["id"]) for t in types},
                      data["agent"].age, data["agent"].sex, 16, 1024, DEV, orig_id=data["agent"]["original_index"])
    assert ref.type_tier[types.index("household")] == W.TIER_RANGE and ref.type_tier[types.index("leisure")] == W.TIER_CELL

    class OnHost:   # _compare reads tensors: bring the native arrays to the CPU
        def __init__(self, nw):
            self.nw = nw

        def desc(self):
            return self.nw.desc()

        def array(self, *a, **k):
            return self.nw.array(*a, **k).cpu()

    _compare(OnHost(native), ref, types)


def test_c_abi_only_build_and_step_of_the_sample_world(golden_dir):
    from grad_june import _lib                      # the ctypes binding of include/gradjune_b200.h, nothing else
    from grad_june.default_config import default_parameters
    L = _lib.lib()
    arrays = np.load(golden_dir / "sample_world.npz")
    g = np.load(golden_dir / "run_sample_default.npz")
    types = H.SAMPLE_TYPES
    n = len(arrays["age"])
    keep = []

    def dev(x, dtype):
        t = (x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))).to(device=DEV, dtype=dtype).contiguous()
        keep.append(t)
        return t

    # ---- gj_world_build from the reference's arrays ---------------------------------------------------------
    src = _lib.WorldSrc()
    src.n_agents, src.n_types, src.renumber = n, len(types), 1
    for i, t in enumerate(types):
        src.type_name[i] = t.encode()
        src.edge_agent[i] = dev(arrays[f"{t}_src"], torch.long).data_ptr()
        src.edge_group[i] = dev(arrays[f"{t}_dst"], torch.long).data_ptr()
        src.n_edges[i], src.n_groups[i] = len(arrays[f"{t}_src"]), int(arrays[f"{t}_ngroups"])
        src.people_i64[i] = dev(arrays[f"{t}_people"], torch.long).data_ptr()
    src.age, src.sex = dev(arrays["age"], torch.long).data_ptr(), dev(arrays["sex"], torch.long).data_ptr()
    world = C.c_void_p()
    assert L.gj_world_build(C.byref(src), C.byref(world)) == 0, L.gj_world_last_error()
    desc = L.gj_world_descriptor(world)
    assert desc.contents.n_agents == n and list(desc.contents.type_tier[:6]) == [1, 0, 0, 0, 0, 2]
    perm = torch.empty(n, dtype=torch.long, device=DEV)
    p = L.gj_world_permutation(world)
    assert p
    assert L.gj_memcpy(perm.data_ptr(), p, n * 8) == 0

    # ---- state + profile in the LOADED numbering, gathered through the permutation -----------------------------
    now, dt = 9.0, 1.0
    state = H.mid_epidemic_state(n, now, 4, DEV)
    prof = {k: v.to(DEV) for k, v in H.profile_params(g).items()}
    lay = lambda t: t[perm].contiguous()   # noqa: E731
    st = L.gj_step_forward.argtypes and torch.cuda.current_stream().cuda_stream
    maxinf, shape, rate, shift = (lay(prof[k]) for k in ("max_infectiousness", "shape", "rate", "shift"))
    k0, prof4 = torch.empty(n, device=DEV), torch.empty(n, 4, device=DEV)
    assert L.gj_profile_prepare(n, shape.data_ptr(), k0.data_ptr(), st) == 0
    assert L.gj_profile_pack(n, maxinf.data_ptr(), shape.data_ptr(), rate.data_ptr(), shift.data_ptr(), k0.data_ptr(),
                             prof4.data_ptr(), st) == 0
    params = default_parameters()
    from grad_june.symptoms import SymptomsSampler   # host-side table parsing of the YAML (no kernels)
    from grad_june.infection_networks import InfectionNetworks
    sampler = SymptomsSampler.from_parameters({**params, "system": {"device": "cpu"}})
    tabs = sampler.tables("cpu")
    nets = InfectionNetworks.from_parameters({**params, "system": {"device": "cpu"}})
    order = ["school", "university", "company", "care_home", "pub", "gym", "grocery", "visit", "care_visit", "cinema",
             "household"]
    P = _lib.StepParams()
    P.mode, P.phases, P.now, P.dt, P.day_type = 0, 15, now, dt, 0
    P.n_nets = len(order)
    leisure_rows, tables, off = {}, [], 0
    G = {t: int(arrays[f"{t}_ngroups"]) for t in types}
    for k, name in enumerate(order):
        net = nets.networks[name]
        ti = types.index(net.edge_type())
        P.nets[k].type, P.nets[k].kind, P.nets[k].s_off = ti, net.kind, off
        P.nets[k].prob_row = -1
        if net.kind in (2, 3):
            P.nets[k].prob_row = len(tables)
            tables.append(net.leisure_probabilities.float())
        off += G[net.edge_type()]
    P.n_quar, P.n_stages, P.tau = -1, tabs.n_stages, 0.1
    for i in range(_lib.GJ_MAX_STAGES):
        for arr, table in ((P.trans_time, tabs.trans), (P.rec_time, tabs.rec)):
            e = table.get(i)
            arr[i].kind = -1 if e is None else int(e[0])
            if e is not None:
                arr[i].loc, arr[i].scale = float(e[1]), float(e[2])
    P.n_age_bins = 3
    for i, b in enumerate((0, 18, 65, 100)):
        P.age_bins[i] = b
    P.seed, P.call_index = 777, 0
    beta = dev([10.0 ** float(nets.networks[k].log_beta) for k in order], torch.float32)
    io = _lib.FwdIO()
    ins = {"s": "susceptibility", "inf": "is_infected", "tinf": "infection_time", "cur": "current_stage",
           "nxt": "next_stage", "ttn": "time_to_next_stage"}
    for k, name in ins.items():
        setattr(io, k, dev(lay(state[name]), torch.float32).data_ptr())
    outs = {}
    for k in ("s_o", "inf_o", "tinf_o", "cur_o", "nxt_o", "ttn_o", "T", "q", "n", "tape_v", "tape_y0"):
        outs[k] = torch.empty(n, device=DEV)
        setattr(io, k, outs[k].data_ptr())
    io.Tq = outs["T"].data_ptr()
    for k, t in (("maxinf", maxinf), ("shape", shape), ("rate", rate), ("shift", shift), ("k0", k0), ("prof4", prof4),
                 ("beta", beta), ("leisure_prob", dev(torch.stack(tables), torch.float32)),
                 ("stage_prob", dev(tabs.stage_prob, torch.float32))):
        setattr(io, k, t.data_ptr())
    n_sum = off + desc.contents.n_groups
    S_sc, S_un, red = torch.empty(n_sum, device=DEV), torch.empty(n_sum, device=DEV), torch.empty(5, device=DEV)
    scratch = torch.zeros(L.gj_scratch_bytes(desc), dtype=torch.uint8, device=DEV)
    io.S_scaled, io.S_unscaled, io.red, io.scratch = S_sc.data_ptr(), S_un.data_ptr(), red.data_ptr(), scratch.data_ptr()
    assert L.gj_step_plan(desc, C.byref(P), None, 0) == 1          # the throughput kernels
    assert L.gj_step_forward(desc, C.byref(P), C.byref(io), st) == 0, L.gj_last_error()
    torch.cuda.synchronize()

    # ---- the oracle on the world as loaded, the kernels' own noise ---------------------------------------------
    E, u, z = (torch.empty(2, n, device=DEV), torch.empty(n, device=DEV), torch.empty(n, device=DEV))
    assert L.gj_philox_fill(777, 0, n, E.data_ptr(), u.data_ptr(), z.data_ptr(), st) == 0
    w = H.oracle_world(arrays, types)
    spec = O.StepSpec(now=now, dt=dt, day_type=0, quarantine=None,
                      nets=[O.NetSpec(k, nets.networks[k].edge_type(), nets.networks[k].kind,
                                      10.0 ** nets.networks[k].log_beta, getattr(nets.networks[k], "leisure_probabilities", None))
                            for k in order])
    ost = {k: v.cpu().clone() for k, v in state.items()}
    aux = {}
    O.step(w, ost, {k: v.cpu() for k, v in prof.items()}, spec, H.oracle_symptoms(sampler),
           O.StepNoise(E=E.cpu(), u=u.cpu(), z=z.cpu().expand(10, n)), aux)
    back = torch.empty_like(perm)
    back[perm] = torch.arange(n, device=DEV)
    orig = lambda t: t[back].cpu().numpy()   # noqa: E731
    q, qo = orig(outs["q"]), aux["q"].numpy()
    assert np.max(np.abs(q - qo) / qo) <= 1e-5
    assert np.array_equal(orig(outs["n"]), aux["new_infected"].numpy()) and aux["new_infected"].sum() > 0
    for k, name in (("inf_o", "is_infected"), ("s_o", "susceptibility"), ("cur_o", "current_stage"), ("nxt_o", "next_stage")):
        assert np.array_equal(orig(outs[k]), ost[name].numpy()), name
    assert red[0].item() == ost["is_infected"].sum().item()
    assert L.gj_world_destroy(world) == 0


def test_runner_on_the_native_world_is_bit_identical(monkeypatch):
    """The whole drop-in path (Runner on the renumbered sample world, Philox mode, backward) with the world built by
    gj_world_build instead of the torch builder: identical arrays, so bit-identical results and gradients."""
    from grad_june import ops
    from grad_june.world import NativeWorld, get_device_world
    from gpu_helpers import make_runner
    outs = []
    for native in ("0", "1"):
        monkeypatch.setenv("GJ_NATIVE_BUILD", native)
        runner, g, params = make_runner("sample_policies", DEV, renumber=True)
        assert isinstance(get_device_world(runner.data, DEV), NativeWorld) == (native == "1")
        with ops.philox_seed(5):
            results, is_inf = runner()
        (results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum()).backward()
        nets = runner.model.infection_networks.networks
        outs.append((results["cases_per_timestep"].detach().clone(), is_inf.detach().clone(),
                     torch.stack([nets[k].log_beta.grad for k in nets.keys()]).clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert outs[0][0][-1] > outs[0][0][0]
