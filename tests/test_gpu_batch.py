"""Batched ensemble [N, b] (BASELINE config 5): b parameter samples stepped by ONE call of
gj_step_forward_batch / gj_step_backward_batch on a shared world.  Sample r of a batched run must be the unbatched
run with log_beta[r]: same Philox stream, same kernels' arithmetic — trajectories bit-identical, gradients equal up
to the order in which per-CTA partial sums of d/dbeta are combined (the batched grid gives every sample fewer CTAs)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _runner(n_agents, policies, days=4, seed=8):
    from grad_june import GradJune, Timer
    from grad_june.default_config import default_parameters
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    params = default_parameters()
    params["system"]["device"] = DEV
    params["timer"]["total_days"] = days
    params["infection_seed"]["log_fraction_initial_cases"] = -1.5
    if policies:
        params["policies"] = {"quarantine": {"quarantine": {1: {"start_date": "2022-02-03", "end_date": "2023-01-01",
                                                                 "stage_threshold": 4}}}}
    torch.manual_seed(5)
    data = Runner.get_data(params, data=make_synthetic_world(n_agents, seed=seed, device=DEV, agents_per_super_area=5000))
    model = GradJune.from_parameters(params)
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test_batch", parameters=params)
    return runner, params, list(model.infection_networks.networks.keys())


def _loss(r):
    # one loss per sample for [T+1, b] series, a scalar for [T+1] series
    return (r["cases_per_timestep"].sum(0) + 3.0 * r["deaths_per_timestep"].sum(0) + 0.5 * r["cases_by_age_65"].sum(0)
            + 0.25 * r["daily_cases_per_timestep"][-1])


def _samples(params, keys, b):
    base = torch.tensor([float(params["networks"][k]["log_beta"]) + 0.4 for k in keys], device=DEV)
    return torch.stack([base + (r - 1) * torch.linspace(-0.25, 0.25, len(keys), device=DEV) for r in range(b)])


def _close(a, b, rtol):
    return torch.allclose(a, b, rtol=rtol, atol=rtol * 1e-2 * float(b.abs().max()))


@pytest.mark.parametrize("policies,n_agents", [(False, 120_002), (True, 90_001)])
def test_batched_window_matches_unbatched_samples(policies, n_agents):
    """Eager Runner with runner.batch = 3 (n_agents not a multiple of 4: padded rows) against three unbatched runs:
    result series, final is_infected and states bit-identical; d loss / d log_beta and d loss / d log_fraction to
    2e-6 (combination order of per-CTA partials only)."""
    from grad_june import ops
    runner, params, keys = _runner(n_agents, policies)
    model = runner.model
    nets = model.infection_networks.networks
    b = 3
    lbs = _samples(params, keys, b)
    assert model.kernel_family(runner.data, runner.timer) == "throughput"

    def run(batch, lb):
        runner.batch = batch
        leaves = []
        for i, k in enumerate(keys):
            leaf = (lb[:, i] if batch else lb[i]).detach().clone().requires_grad_(True)
            nets[k].log_beta = leaf
            leaves.append(leaf)
        frac = torch.tensor(-1.5, device=DEV, requires_grad=True)
        runner.log_fraction_initial_cases = frac
        with ops.philox_seed(321):
            results, is_inf = runner()
        loss = _loss(results)
        loss.sum().backward()
        sym = runner.data["agent"].symptoms
        n = runner.n_agents
        state = [runner.data["agent"].susceptibility, runner.data["agent"].infection_time, sym["current_stage"],
                 sym["next_stage"], sym["time_to_next_stage"]]
        state = [t.detach()[..., :n].clone() for t in state]
        series = {k: v.detach().clone() for k, v in results.items() if k != "dates"}
        grads = torch.stack([l.grad for l in leaves], dim=-1)        # [K] or [b, K]
        return loss.detach().clone(), grads, frac.grad.clone(), series, is_inf.detach().clone(), state

    singles = [run(None, lbs[r]) for r in range(b)]
    loss_b, grads_b, gfrac_b, series_b, inf_b, state_b = run(b, lbs)
    assert tuple(loss_b.shape) == (b,) and tuple(grads_b.shape) == (b, len(keys)) and tuple(inf_b.shape) == (b, n_agents)
    worst = 0.0
    for r in range(b):
        loss_s, grads_s, _, series_s, inf_s, state_s = singles[r]
        for k, v in series_s.items():
            assert torch.equal(series_b[k][:, r], v), (k, r)
        assert torch.equal(inf_b[r], inf_s), r
        for tb, ts in zip(state_b, state_s):
            assert torch.equal(tb[r], ts), r
        assert torch.equal(loss_b[r], loss_s)
        assert _close(grads_b[r], grads_s, 2e-6), (r, grads_b[r], grads_s)
        worst = max(worst, float(((grads_b[r] - grads_s).abs() / grads_s.abs().clamp_min(1e-30)).max()))
    gfrac_sum = sum(s[2] for s in singles)
    assert _close(gfrac_b, gfrac_sum, 2e-6), (gfrac_b, gfrac_sum)
    # the samples really differ, and the epidemic really runs
    assert not torch.equal(series_b["cases_per_timestep"][:, 0], series_b["cases_per_timestep"][:, 2])
    assert series_b["cases_per_timestep"][-1, 1] > series_b["cases_per_timestep"][0, 1] > 0
    print(f"batched vs unbatched: largest relative gradient deviation {worst:.2e}")
    runner.batch = None


def test_batched_graph_replay_and_ensemble():
    """GraphedRunner(batch=4): one graph replay steps four samples; every row equals the unbatched replay with that
    row's log-betas.  EnsembleEvaluator(batch=2) on five samples (the last replay padded) equals the unbatched one."""
    from grad_june.calibration import EnsembleEvaluator
    from grad_june.graphed import GraphedRunner
    runner, params, keys = _runner(100_000, True, days=3)
    b = 4
    lbs = _samples(params, keys, 5)
    single = GraphedRunner(runner, _loss, seed=77)
    ref = []
    for r in range(5):
        loss, grads, results = single(lbs[r])
        torch.cuda.synchronize()
        ref.append((loss.clone(), grads.clone(), results["cases_per_timestep"].clone(), single.is_infected.clone()))
    batched = GraphedRunner(runner, _loss, seed=77, batch=b)
    for rows in ([0, 1, 2, 3], [4, 2, 0, 1], [0, 1, 2, 3]):
        loss, grads, results = batched(lbs[rows])
        torch.cuda.synchronize()
        for j, r in enumerate(rows):
            assert torch.equal(results["cases_per_timestep"][:, j], ref[r][2]), (rows, j)
            assert torch.equal(batched.is_infected[j], ref[r][3])
            assert torch.equal(loss[j], ref[r][0])
            assert _close(grads[j], ref[r][1], 2e-6), (grads[j], ref[r][1])
    del batched
    ens = EnsembleEvaluator(runner, _loss, seed=77, batch=2)
    losses, grads = ens(lbs)
    assert tuple(losses.shape) == (5,) and tuple(grads.shape) == (5, len(keys))
    for r in range(5):
        assert torch.equal(losses[r], ref[r][0])
        assert _close(grads[r], ref[r][1], 2e-6)
    runner.batch = None


def test_batch_entry_points_reject_what_they_cannot_run():
    """gj_step_forward_batch refuses (error code + message, nothing launched) a stride that is not a multiple of four
    agents, a scratch stride smaller than gj_scratch_bytes, and the seeding mode."""
    from grad_june import _lib, ops
    runner, params, keys = _runner(20_000, False, days=1)
    static, rows = runner.model._static(runner.data, torch.device(DEV))
    runner.timer.reset()
    spec, nets = runner.model._spec(runner.timer, rows, (0, 18, 65, 100), ops.MODE_STEP, False)
    world = static.world
    p, _ = ops._fill_params(world, spec, static.symptoms, 1, 0)
    L = _lib.lib()
    io = _lib.FwdIO()
    dummy = torch.zeros(16, device=DEV)
    io.scratch = dummy.data_ptr()
    scr = int(L.gj_scratch_bytes(C.byref(world.desc())))
    good = dict(n_samples=2, agent_stride=20_000, group_stride=1 << 20, beta_stride=p.n_nets, red_stride=5,
                scratch_stride=(scr + 255) // 256 * 256)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for bad, text in ((dict(agent_stride=20_002), "multiple of 4"), (dict(scratch_stride=256), "scratch_stride"),
                      (dict(n_samples=0), "n_samples")):
        bt = _lib.Batch(**{**good, **bad})
        rc = L.gj_step_forward_batch(C.byref(world.desc()), C.byref(p), C.byref(io), C.byref(bt), stream)
        assert rc < 0 and text in L.gj_last_error().decode(), (rc, L.gj_last_error())
    p.mode = _lib.MODE_SEED
    bt = _lib.Batch(**good)
    rc = L.gj_step_forward_batch(C.byref(world.desc()), C.byref(p), C.byref(io), C.byref(bt), stream)
    assert rc < 0 and "fused step" in L.gj_last_error().decode()
