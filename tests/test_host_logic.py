"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol the header declares,
Philox known answers, the default configuration equals the reference's, Timer/Policies reproduce the
reference's schedule, the world container reads reference pickles' layout, and the CSR/tier builder."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from grad_june import _lib
    header = (ROOT / "include" / "gradjune_b200.h").read_text()
    declared = set(re.findall(r"\b(gj_[a-z0-9_]+)\s*\(", header))
    declared -= {"gj_world_desc", "gj_step_params"}
    assert {"gj_step_forward", "gj_step_backward", "gj_transmission_forward", "gj_philox_fill"} <= declared
    raw = ctypes.CDLL(str(_lib.build()))
    for sym in sorted(declared):
        assert hasattr(raw, sym), f"{sym} is declared in include/gradjune_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) <= declared
    L = _lib.lib()                     # also checks struct sizes against the header's
    assert L.gj_abi_version() == _lib.GJ_ABI_VERSION


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    from grad_june import _lib
    L = _lib.lib()

    def philox(ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        out = (ctypes.c_uint32 * 4)()
        L.gj_philox4x32_10(c, k, out)
        return [int(x) for x in out]

    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]

    # Philox2x32-10 (the step's two exponentials): Random123 kat_vectors
    def philox2(ctr, key):
        c = (ctypes.c_uint32 * 2)(*ctr)
        out = (ctypes.c_uint32 * 2)()
        L.gj_philox2x32_10(c, key, out)
        return list(out)
    assert philox2([0, 0], 0) == [0xff1dae59, 0x6cd10df2]
    assert philox2([0xffffffff, 0xffffffff], 0xffffffff) == [0x2c3f628b, 0xab4fd7ad]
    assert philox2([0x243f6a88, 0x85a308d3], 0x13198a2e) == [0xdd7ce038, 0xf62a4c12]


def test_default_config_equals_reference(golden_dir):
    from grad_june.default_config import default_parameters
    ref = json.load(open(golden_dir / "reference_default_params.json"))
    mine = json.loads(json.dumps(default_parameters(), default=str, sort_keys=True))
    assert mine == ref


def test_default_yaml_roundtrip(tmp_path):
    import yaml
    from grad_june.default_config import default_parameters, write_default_config
    write_default_config(tmp_path / "default.yaml")
    assert yaml.safe_load(open(tmp_path / "default.yaml")) == default_parameters()


@pytest.mark.parametrize("tag", list(H.RUNS))
def test_schedule_matches_reference(tag):
    """Timer + Policies + network table give the reference's per-step (now, dt, day type, activity order,
    beta_eff, active quarantine thresholds) — recorded from the reference by make_golden.py."""
    from grad_june import Timer
    from grad_june.infection_networks import InfectionNetworks
    from grad_june.policies import Policies
    params, schedule = H.load_params(tag)
    nets = InfectionNetworks.from_parameters(params)
    policies = Policies.from_parameters(params)
    timer = Timer.from_parameters(params)
    i = 0
    while timer.date < timer.final_date:
        next(timer)
        ref = schedule[i]
        assert timer.now == ref["now"] and timer.duration == ref["dt"] and timer.day_type == ref["day_type"]
        assert timer.date.isoformat() == ref["date"]
        active = nets.active_networks(timer, policies)
        assert [n.name for n in active] == ref["order"]
        for n in active:
            assert float(n.beta_eff(policies, timer)) == pytest.approx(ref["beta"][n.name], rel=1e-7)
        q = policies.quarantine_policies.active_thresholds(timer) if policies.quarantine_policies else None
        assert q == ref["quarantine"]
        i += 1
    assert i == len(schedule)


def test_timer_shifts_and_unknown_activity():
    from grad_june import Timer
    t = Timer(initial_day="2022-02-04", total_days=3, weekday_step_duration=(8, 8, 8), weekend_step_duration=(12, 12),
              weekday_activities=(("company", "school", "household"), ("pub", "household"), ("household",)),
              weekend_activities=(("pub",), ("household",)))
    assert t.get_activity_order() == ["school", "company", "household"]      # hierarchy order, not config order
    seen = []
    while t.date < t.final_date:
        seen.append((t.day_of_week, t.shift, t.duration))
        next(t)
    assert seen[:3] == [("Friday", 0, 8 / 24), ("Friday", 1, 8 / 24), ("Friday", 2, 8 / 24)]
    assert seen[3:5] == [("Saturday", 0, 0.5), ("Saturday", 1, 0.5)]
    bad = Timer(initial_day="2022-02-01", weekday_activities=(("leisure",),), weekday_step_duration=(24,))
    with pytest.raises(ValueError):
        bad.get_activity_order()


def test_parse_age_probabilities_overlapping_bins():
    from grad_june.utils import parse_age_probabilities
    out = parse_age_probabilities({"0-75": 0.0, "75-85": 0.25, "75-100": 0.5})
    assert len(out) == 100 and out[10] == 0.0 and out[74] == 0.0
    assert out[75] in (0.25, 0.5) and out[99] in (0.0, 0.5)
    simple = parse_age_probabilities({"20-40": 0.3, "0-20": 0.1})
    assert simple[0] == 0.1 and simple[19] == 0.1 and simple[20] == 0.3 and simple[39] == 0.3 and simple[40] == 0


def test_world_container_and_pickle_roundtrip(tmp_path, golden_dir):
    import pickle
    from grad_june.world import HeteroData, ToUndirected, load_world, world_from_arrays
    arrays = np.load(golden_dir / "sample_world.npz")
    data = world_from_arrays(arrays, H.SAMPLE_TYPES)
    assert data["agent"].age.shape[0] == 769 and data["agent"]["age"] is data["agent"].age
    assert data["attends_school"].edge_index.shape[0] == 2
    assert torch.equal(data["rev_attends_school"].edge_index, data["attends_school"].edge_index.flip(0))
    assert data["school"]["people"].shape[0] == len(data["school"]["id"])
    data["results"] = {"x": 1}
    assert data["results"]["x"] == 1 and data.results["x"] == 1
    del data["rev_attends_school"]
    assert "rev_attends_school" not in data
    data = ToUndirected()(data)
    with open(tmp_path / "w.pkl", "wb") as f:
        pickle.dump(data, f)
    again = load_world(tmp_path / "w.pkl")
    assert torch.equal(again["attends_leisure"].edge_index, data["attends_leisure"].edge_index)
    assert set(again.venue_types()) == set(H.SAMPLE_TYPES)


def _group_sums_reference(ei, n_groups, values):
    return torch.zeros(n_groups, dtype=torch.float64).index_add_(0, ei[1], values.double()[ei[0]])


def test_world_tiers_reproduce_group_sums():
    """Whatever layout tier a type lands in, the stored structure must describe the same agent<->group
    incidence: recompute per-group sums of a random per-agent vector from each tier and compare."""
    from grad_june import world as W
    n = 40_000
    data = W.make_synthetic_world(n, seed=4, agents_per_super_area=3000)
    types = data.venue_types()
    dw = W.build_csr(n, types, {t: data["attends_" + t].edge_index for t in types},
                     {t: data[t]["people"] for t in types}, {t: len(data[t]["id"]) for t in types},
                     data["agent"].age, data["agent"].sex, 16, 1024, "cpu")
    tiers = dict(zip(types, dw.type_tier))
    assert tiers["household"] == W.TIER_RANGE and tiers["leisure"] == W.TIER_CELL and tiers["company"] == W.TIER_GENERIC
    x = torch.rand(n, dtype=torch.float64)
    u32 = lambda t: t.long() & 0xFFFFFFFF
    tile_begin = u32(dw.tile_begin)
    assert tile_begin[0] == 0 and tile_begin[-1] == n and int((tile_begin[1:] - tile_begin[:-1]).max()) <= W.TILE_AGENTS
    from grad_june import _lib
    assert W.TILE_AGENTS == _lib.config()["tile_agents"]      # the pipelined kernels stage one tile per thread block
    for ti, t in enumerate(types):
        ei = data["attends_" + t].edge_index
        G = len(data[t]["id"])
        ref = _group_sums_reference(ei, G, x)
        if dw.type_tier[ti] == W.TIER_RANGE:
            slot = u32(dw.range_slot[ti])
            member = slot != 0xFFFFFFFF
            start = torch.arange(n)[member] - (slot[member] >> 16)
            size = slot[member] & 0xFFFF
            # every member sees the same (start, size); group sum = sum over the run
            csum = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(x, 0)))
            mine = csum[start + size] - csum[start]
            g_of = torch.zeros(n, dtype=torch.long)
            g_of[ei[0]] = ei[1]
            assert torch.allclose(mine, ref[g_of[member]], rtol=1e-9, atol=1e-9)
            pc = W.p_contact(data[t]["people"])
            assert torch.equal(dw.range_pc[ti][member], pc[g_of[member]])
        elif dw.type_tier[ti] == W.TIER_CELL:
            c = dw.cells[ti]
            ctp, cgp, cg = u32(c["cell_tile_ptr"]), u32(c["cell_grp_ptr"]), u32(c["cell_grp"])
            mine = torch.zeros(G, dtype=torch.float64)
            csum = torch.cat((torch.zeros(1, dtype=torch.float64), torch.cumsum(x, 0)))
            for cell in range(c["n_cells"]):
                lo, hi = tile_begin[ctp[cell]], tile_begin[ctp[cell + 1]]
                for g in cg[cgp[cell]:cgp[cell + 1]]:
                    mine[g] += csum[hi] - csum[lo]
            assert torch.allclose(mine, ref, rtol=1e-9, atol=1e-9)
            # reverse map is consistent
            gcp, gc = u32(c["grp_cell_ptr"]), u32(c["grp_cell"])
            assert int(gcp[-1]) == cg.numel() == gc.numel()
        else:
            off = dw.type_group_off[ti]
            gm_ptr, gm_agent = u32(dw.gm_ptr), u32(dw.gm_agent)
            seg = torch.repeat_interleave(torch.arange(G), (gm_ptr[off + 1:off + G + 1] - gm_ptr[off:off + G]))
            vals = x[gm_agent[gm_ptr[off]:gm_ptr[off + G]]]
            mine = torch.zeros(G, dtype=torch.float64).index_add_(0, seg, vals)
            assert torch.allclose(mine, ref, rtol=1e-9, atol=1e-9)
    # work lists cover every generic group exactly once
    small, chunks = u32(dw.small_groups), u32(dw.chunk_group)
    generic = torch.cat([torch.arange(dw.type_group_off[i], dw.type_group_off[i + 1])
                         for i in range(len(types)) if dw.type_tier[i] == W.TIER_GENERIC])
    assert torch.equal(torch.sort(torch.cat((small, torch.unique(chunks))))[0], generic)


def test_untiered_world_matches_tiered_incidence():
    from grad_june import world as W
    data = W.make_synthetic_world(5000, seed=2, agents_per_super_area=1000)
    types = data.venue_types()
    args = (5000, types, {t: data["attends_" + t].edge_index for t in types}, {t: data[t]["people"] for t in types},
            {t: len(data[t]["id"]) for t in types}, data["agent"].age, data["agent"].sex, 16, 1024, "cpu")
    flat = W.build_csr(*args, tiers=False)
    assert set(flat.type_tier) == {W.TIER_GENERIC} and flat.n_generic_edges == flat.n_edges
    tiered = W.build_csr(*args)
    assert tiered.n_generic_edges < flat.n_edges and tiered.n_edges == flat.n_edges


def test_cpu_tensors_fail_loudly():
    """There is no CPU fallback: the modules refuse CPU tensors instead of silently computing elsewhere."""
    from grad_june import IsInfectedSampler, _lib
    with pytest.raises(_lib.GradJuneLibraryError):
        IsInfectedSampler()(torch.full((8,), 0.5))


def _build(data, orig=None):
    from grad_june import world as W
    types = data.venue_types()
    n = len(data["agent"].id)
    return W.build_csr(n, types, {t: data["attends_" + t].edge_index for t in types},
                       {t: torch.as_tensor(data[t]["people"]) for t in types}, {t: len(data[t]["id"]) for t in types},
                       data["agent"].age, data["agent"].sex, 16, 1024, "cpu",
                       orig_id=data["agent"]["original_index"] if "original_index" in data["agent"] else None)


def test_renumbering_makes_the_reference_sample_world_streamable(golden_dir):
    """The reference's sample world as loaded (agents by area and age: households scattered) has its household on
    the GENERIC tier; world.renumber_world puts it on the RANGE tier and leisure on the CELL tier without changing
    the agent<->group incidence, and original_order / layout_order_of map per-agent values between numberings."""
    from grad_june import world as W
    arrays = np.load(golden_dir / "sample_world.npz")
    data = W.world_from_arrays(arrays, H.SAMPLE_TYPES)
    before = dict(zip(data.venue_types(), _build(data).type_tier))
    assert before["household"] == W.TIER_GENERIC and before["leisure"] == W.TIER_CELL
    n = len(data["agent"].id)
    age0 = data["agent"].age.clone()
    edges0 = {t: data["attends_" + t].edge_index.clone() for t in data.venue_types()}
    data = W.renumber_world(data)
    oi = data["agent"].original_index
    assert torch.equal(torch.sort(oi)[0], torch.arange(n))
    assert torch.equal(data["agent"].age, age0[oi])
    assert torch.equal(W.original_order(data, data["agent"].age), age0)
    assert torch.equal(W.layout_order_of(data, age0), data["agent"].age)
    dw = _build(data)
    after = dict(zip(data.venue_types(), dw.type_tier))
    assert after["household"] == W.TIER_RANGE and after["leisure"] == W.TIER_CELL
    assert all(after[t] == W.TIER_GENERIC for t in ("company", "school", "university", "care_home"))
    assert torch.equal(dw.orig_id.long() & 0xFFFFFFFF, oi)
    for t, e0 in edges0.items():          # same edges, same order, agents renamed
        e1 = data["attends_" + t].edge_index
        assert torch.equal(oi[e1[0]], e0[0]) and torch.equal(e1[1], e0[1])
        assert torch.equal(data["rev_attends_" + t].edge_index, e1.flip(0))
    # household members: consecutive ids, in the reference's edge order
    hh = data["attends_household"].edge_index
    _, order = torch.sort(hh[1], stable=True)
    m, gidx = hh[0][order], hh[1][order]
    same = gidx[1:] == gidx[:-1]
    assert bool((m[1:][same] == m[:-1][same] + 1).all())
    # a second renumbering finds nothing to do
    assert W.layout_order(n, data.venue_types(), {t: data["attends_" + t].edge_index for t in data.venue_types()},
                          {t: len(data[t]["id"]) for t in data.venue_types()}) is None


def test_renumbering_recovers_a_shuffled_synthetic_world():
    from grad_june import world as W
    n = 30_000
    data = W.make_synthetic_world(n, seed=3, agents_per_super_area=2500)
    assert W.renumber_world(data) is data and "original_index" not in data["agent"]     # already laid out
    ref = _build(data)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
    data = W.renumber_world(data, perm)
    assert torch.equal(data["agent"].original_index, perm)
    shuffled = dict(zip(data.venue_types(), _build(data).type_tier))
    assert shuffled["household"] == W.TIER_GENERIC and shuffled["leisure"] == W.TIER_GENERIC
    data = W.renumber_world(data)
    # original_index composes: it still refers to the first numbering, and the layout is the first one again
    assert torch.equal(data["agent"].original_index, torch.arange(n))
    dw = _build(data)
    assert dw.type_tier == ref.type_tier
    for name in ("tile_begin", "ent1", "gm_agent", "gm_ptr", "am_ent", "cls"):
        assert torch.equal(getattr(dw, name), getattr(ref, name)), name
    ti = data.venue_types().index("household")
    assert torch.equal(dw.range_slot[ti], ref.range_slot[ti])


def test_tier_policy_one_range_type_and_leisure_cells_only():
    """At most one RANGE-tier type (the throughput kernels re-sum one household-like network per agent) and the
    CELL tier for "leisure" only, whatever other types happen to look like (university members of the sample
    world are contiguous runs: they stay GENERIC)."""
    from grad_june import world as W
    n = 4000
    data = W.make_synthetic_world(n, seed=5, agents_per_super_area=1000)
    ids = torch.arange(n)
    data["pair"].id = torch.arange(n // 2)
    data["pair"].people = torch.full((n // 2,), 2)
    data["agent", "attends_pair", "pair"].edge_index = torch.stack((ids, ids // 2))      # also household-like
    data["block"].id = torch.arange(4)
    data["block"].people = torch.full((4,), n // 4)
    data["agent", "attends_block", "block"].edge_index = torch.stack((ids, ids // (n // 4)))   # also cell-like
    tiers = dict(zip(data.venue_types(), _build(data).type_tier))
    assert tiers["household"] == W.TIER_RANGE and tiers["pair"] == W.TIER_GENERIC
    assert tiers["leisure"] == W.TIER_CELL and tiers["block"] == W.TIER_GENERIC


def test_batched_ensemble_host_logic():
    """Host side of the batched ensemble (no GPU): beta_vector turns per-network [b] log-betas into a [b, K] beta table
    (uncalibrated scalars broadcast), differentiable per sample; the gj_batch binding has the library's layout; a
    batched window on a CPU world is refused loudly."""
    from grad_june import _lib
    from grad_june.infection_networks import InfectionNetworks
    from grad_june.infection_networks.base import beta_vector
    from grad_june.default_config import default_parameters

    params = default_parameters()
    nets_mod = InfectionNetworks.from_parameters(params)
    nets = list(nets_mod.networks.values())[:4]
    b = 3
    lb = torch.linspace(-0.3, 0.2, b * 2).reshape(b, 2).requires_grad_(True)
    nets[0].log_beta = lb[:, 0]
    nets[2].log_beta = lb[:, 1]
    fixed1, fixed3 = float(nets[1].log_beta), float(nets[3].log_beta)
    beta = beta_vector(nets, None, None, "cpu")
    assert tuple(beta.shape) == (b, 4) and beta.is_contiguous()
    assert torch.allclose(beta[:, 0], 10.0 ** lb[:, 0]) and torch.allclose(beta[:, 2], 10.0 ** lb[:, 1])
    assert torch.allclose(beta[:, 1], torch.full((b,), 10.0 ** fixed1)) and torch.allclose(beta[:, 3], torch.full((b,), 10.0 ** fixed3))
    (beta * torch.arange(1.0, b + 1).reshape(b, 1)).sum().backward()
    expect = torch.log(torch.tensor(10.0)) * (10.0 ** lb.detach()) * torch.arange(1.0, b + 1).reshape(b, 1)
    assert torch.allclose(lb.grad, expect, rtol=1e-6)
    # unbatched call unchanged: [K]
    for n in nets:
        n.log_beta = torch.tensor(0.25)
    assert tuple(beta_vector(nets, None, None, "cpu").shape) == (4,)
    # binding layout = library layout (checked at load time as well)
    cfg = (ctypes.c_int64 * 10)()
    assert _lib.lib().gj_config(cfg, 10) == 10 and cfg[9] == ctypes.sizeof(_lib.Batch)
    for sym in ("gj_step_forward_batch", "gj_step_backward_batch"):
        assert hasattr(_lib.lib(), sym)


@pytest.mark.parametrize("n_samples,world_size,batch", [(5, 1, 2), (1024, 8, 8), (13, 4, 3), (7, 8, None), (16, 2, 1)])
def test_ensemble_deal_covers_every_sample_once(n_samples, world_size, batch):
    """EnsembleEvaluator's dealing of samples to ranks and replays: every sample evaluated exactly once, at most
    ``batch`` per replay, and the gathered table's row (j, rank) is sample j * world_size + rank."""
    from grad_june.calibration import ensemble_deal
    seen = []
    for rank in range(world_size):
        per, replays = ensemble_deal(n_samples, world_size, rank, batch)
        assert per == -(-n_samples // world_size)
        flat = [i for chunk in replays for i in chunk]
        assert all(1 <= len(chunk) <= (batch or 1) for chunk in replays)
        assert flat == [j * world_size + rank for j in range(len(flat))]      # row j of this rank
        assert len(flat) <= per
        seen += flat
    assert sorted(seen) == list(range(n_samples))


def test_batched_window_start_and_cpu_refusal():
    """Runner.batch = b: the window starts from b copies of the backed-up state, rows padded to a multiple of four
    agents (zeros), cached per batch size; a batched window on CPU tensors is refused (no CPU fallback)."""
    from grad_june import GradJune, Timer
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    params = default_parameters()
    params["system"]["device"] = "cpu"
    params["timer"]["total_days"] = 2
    torch.manual_seed(3)
    data = Runner.get_data(params, data=make_synthetic_world(4001, seed=3, device="cpu", agents_per_super_area=1000))
    n = len(data["agent"].id)
    runner = Runner(model=GradJune.from_parameters(params), data=data, timer=Timer.from_parameters(params),
                    log_fraction_initial_cases=-1.5, save_path="/tmp/gj_test_host", parameters=params)
    runner.data_backup["infection_time"][5] = 2.5
    runner.batch = 3
    runner._restore_for_window()
    agent = runner.data["agent"]
    n_pad = (n + 3) // 4 * 4
    for t in (agent.susceptibility, agent.is_infected, agent.infection_time, agent.symptoms["current_stage"],
              agent.symptoms["next_stage"], agent.symptoms["time_to_next_stage"]):
        assert tuple(t.shape) == (3, n_pad) and t.dtype == torch.float32 and t.is_contiguous()
        assert torch.equal(t[0], t[2]) and float(t[:, n:].abs().sum()) == 0.0
    assert torch.equal(agent.infection_time[1, :n], runner.data_backup["infection_time"])
    assert torch.equal(agent.susceptibility[2, :n], torch.ones(n))
    first = agent.susceptibility
    runner._restore_for_window()
    assert runner.data["agent"].susceptibility is first          # cached
    runner.batch = 2
    runner._restore_for_window()
    assert tuple(runner.data["agent"].susceptibility.shape) == (2, n_pad)
    runner.restore_initial_data()
    with pytest.raises(RuntimeError, match="CUDA"):
        runner()
    runner.batch = None
    runner.restore_initial_data()
    assert tuple(runner.data["agent"].susceptibility.shape) == (n,)
