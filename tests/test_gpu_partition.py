"""Geographic partition on real GPUs (needs >= 2): a world split over two ranks, stepped through Runner() +
backward() with the boundary-group all-reduce over NCCL, against the same world on one GPU with the same Philox
stream (the counter is the global agent id)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N_AGENTS = 200_000
STEPS = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(data, params, dev, exact):
    from grad_june import GradJune, Timer, ops
    from grad_june.runner import Runner
    model = GradJune.from_parameters(params)
    keys = list(model.infection_networks.networks.keys())
    leaves = []
    for k in keys:
        leaf = torch.tensor(float(params["networks"][k]["log_beta"]) + 0.4, device=dev, requires_grad=True)
        model.infection_networks.networks[k].log_beta = leaf
        leaves.append(leaf)
    runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                    save_path="/tmp/gj_test", parameters=params)
    ops.EXACT_ORDER = exact
    try:
        with ops.philox_seed(4242):
            results, is_inf = runner()
        loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum() \
            + 0.5 * results["cases_by_age_65"].sum()
        loss.backward()
    finally:
        ops.EXACT_ORDER = False
    return (results["cases_per_timestep"].detach().cpu().numpy(), is_inf.detach().cpu().numpy(),
            torch.stack([l.grad for l in leaves]))


def _worker(rank, world_size, port, exact, out):
    import torch.distributed as dist
    from grad_june.default_config import default_parameters
    from grad_june.partition import partition_world
    from grad_june.runner import Runner
    from grad_june.world import make_synthetic_world
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device(dev))
    try:
        params = default_parameters()
        params["system"]["device"] = dev
        params["timer"]["total_days"] = STEPS
        params["infection_seed"]["log_fraction_initial_cases"] = -1.5
        params["policies"] = {"quarantine": {"quarantine": {1: {"start_date": "2022-01-01", "end_date": "2023-01-01",
                                                                 "stage_threshold": 4}}}}
        torch.manual_seed(5)
        full = Runner.get_data(params, data=make_synthetic_world(N_AGENTS, seed=5, device=dev,
                                                                 agents_per_super_area=5000))
        local = partition_world(full, rank, world_size)
        part = local._gj_partition
        assert sum(part.n_boundary.values()) > 0
        cases, inf, grads = _run(local, params, dev, exact)
        dist.all_reduce(grads)
        if rank == 0:
            out["part"] = (cases, grads.cpu().numpy())
        out[f"inf{rank}"] = (part.agent_lo, part.agent_hi, inf)
        if rank == 0:
            cases1, inf1, grads1 = _run(full, params, dev, exact)
            out["full"] = (cases1, inf1, grads1.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exact", [False, True])
def test_partitioned_world_matches_single_gpu(exact):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    world_size = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world_size, _free_port(), exact, out), nprocs=world_size, join=True)
    cases1, inf1, grads1 = out["full"]
    cases, grads = out["part"]
    inf = np.empty_like(inf1)
    for r in range(world_size):
        lo, hi, x = out[f"inf{r}"]
        inf[lo:hi] = x
    flips = int((inf != inf1).sum())
    assert cases1[-1] > cases1[0] > 0
    assert flips <= 10, flips          # partial sums are re-associated across ranks: only near-ties may differ
    if flips == 0:
        assert np.array_equal(cases, cases1)
        assert np.allclose(grads, grads1, rtol=1e-4, atol=1e-6 * np.abs(grads1).max()), (grads, grads1)
