"""Host logic of the geographic partition on CPU with the gloo backend (world_size 2 and 3): agent ranges never
split a household, every rank's local world is consistent with the complete one, the packed all-reduce of the
boundary groups reproduces the complete world's group sums on every rank, and every group is owned exactly once."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grad_june import _lib
from grad_june.partition import BoundaryExchange, all_reduce_sum, partition_bounds, partition_world
from grad_june.world import build_csr, make_synthetic_world

N_AGENTS = 60_000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_csr(local):
    cfg = {"small_group": 16, "chunk": 1024}
    types = local.venue_types()
    return build_csr(len(local["agent"].id), types, {t: local["attends_" + t].edge_index for t in types},
                     {t: local[t]["people"] for t in types}, {t: len(local[t]["id"]) for t in types},
                     local["agent"].age, local["agent"].sex, cfg["small_group"], cfg["chunk"], "cpu")


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        data = make_synthetic_world(N_AGENTS, seed=3, device="cpu", agents_per_super_area=3000)
        n = N_AGENTS
        g = torch.Generator().manual_seed(11)
        x = torch.rand(n, generator=g)                      # a per-agent value, the same on every rank
        data["agent"].x = x
        local = partition_world(data, rank, world_size)
        part = local._gj_partition
        lo, hi = part.agent_lo, part.agent_hi
        assert torch.equal(local["agent"].x, x[lo:hi]) and torch.equal(local["agent"].age, data["agent"].age[lo:hi])
        # no household straddles a cut
        hh = data["attends_household"].edge_index
        gid = torch.full((n,), -1, dtype=torch.long)
        gid[hh[0]] = hh[1]
        for c in part.bounds[1:-1]:
            assert gid[c] != gid[c - 1]
        world = _local_csr(local)
        ex = BoundaryExchange(part, world)
        # partial sums of x over this rank's members of every group, laid out like the reference-order buffers
        nets = [(ti, world.type_group_off[ti]) for ti in range(len(world.types))]
        buf = torch.zeros(world.n_groups)
        for ti, t in enumerate(world.types):
            ei = local["attends_" + t].edge_index
            buf[world.type_group_off[ti]:world.type_group_off[ti + 1]].index_add_(0, ei[1], local["agent"].x[ei[0]])
        buf2 = 2.0 * buf
        region = ex.regions(False, 0, nets)
        ex.exchange([buf, buf2], region)
        checked = 0
        for ti, t in enumerate(world.types):
            if world.type_tier[ti] == 1:      # range tier (households): rank-local by construction
                assert part.n_boundary[t] == 0
                continue
            ei = data["attends_" + t].edge_index
            full = torch.zeros(len(data[t]["id"])).index_add_(0, ei[1], x[ei[0]])
            mine = full[part.local_groups[t]]
            got = buf[world.type_group_off[ti]:world.type_group_off[ti + 1]]
            assert torch.allclose(got, mine, rtol=1e-5, atol=1e-5), t
            assert torch.allclose(buf2[world.type_group_off[ti]:world.type_group_off[ti + 1]], 2 * mine, rtol=1e-5, atol=1e-5)
            assert torch.equal(local[t]["people"], torch.as_tensor(data[t]["people"])[part.local_groups[t]])
            checked += part.n_boundary[t]
        assert checked > 0                     # the world does have groups that straddle the cut
        # every group of the complete world that anybody attends is owned by exactly one rank
        for t in world.types:
            G = len(data[t]["id"])
            own = torch.zeros(G)
            own[part.local_groups[t]] = part.owned[t].float()
            dist.all_reduce(own)
            attended = torch.zeros(G)
            attended[data["attends_" + t].edge_index[1]] = 1.0
            assert torch.equal(own, attended), t
        # attendance masks (peer-memory exchange): bit r of a boundary group's mask <=> rank r has a member
        for t in world.types:
            ei = data["attends_" + t].edge_index
            G = len(data[t]["id"])
            r_of = torch.bucketize(ei[0], torch.tensor(part.bounds[1:-1]), right=True)
            truth = torch.zeros(G, dtype=torch.long)
            for r in range(world_size):
                has = torch.zeros(G, dtype=torch.bool)
                has[ei[1][r_of == r]] = True
                truth += has.long() << r
            multi = torch.tensor([bin(int(v)).count("1") >= 2 for v in truth])
            assert torch.equal(part.attend[t], truth[multi]), t
            assert part.n_boundary[t] == int(multi.sum())
            mine_bit = (part.attend[t][part.touch_pos[t]] >> rank) & 1
            assert bool(mine_bit.all())
        region5 = ex.regions(False, 0, nets)
        assert region5[4].numel() == region5[2] and int((region5[4] != 0).all()) == 1
        # differentiable sum over ranks of the per-rank result table
        leaf = torch.full((3,), float(rank + 1), requires_grad=True)
        tot = all_reduce_sum(leaf * 2.0, part)
        tot.sum().backward()
        assert torch.allclose(tot, torch.full((3,), 2.0 * sum(range(1, world_size + 1))))
        assert torch.allclose(leaf.grad, torch.full((3,), 2.0))
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_partition_exchange_gloo(world_size):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world_size, port, out), nprocs=world_size, join=True)
    assert sorted(out.keys()) == list(range(world_size))


def test_partition_bounds_cover_all_agents():
    data = make_synthetic_world(20_000, seed=1, device="cpu", agents_per_super_area=2500)
    for p in (1, 2, 4, 8):
        b = partition_bounds(data, p)
        assert b[0] == 0 and b[-1] == 20_000 and all(b[i] <= b[i + 1] for i in range(p))
        sizes = [b[i + 1] - b[i] for i in range(p)]
        assert max(sizes) - min(sizes) <= 16


# ---- block-wise worlds: every rank generates only its own block (weak-scaling runs) -----------------------
N_BLOCK = 24_000


def _assemble(blocks):
    """The complete world of a list of blocks: agents concatenated, households renumbered consecutively."""
    from grad_june.world import HeteroData
    whole = HeteroData()
    n0 = [0]
    for b in blocks:
        n0.append(n0[-1] + len(b["agent"].id))
    whole["agent"].id = torch.arange(n0[-1])
    for k in ("age", "sex"):
        whole["agent"][k] = torch.cat([b["agent"][k] for b in blocks])
    whole["agent"].ethnicity = blocks[0]["agent"].ethnicity
    for t in blocks[0].venue_types():
        srcs, dsts, g0 = [], [], 0
        for i, b in enumerate(blocks):
            ei = b["attends_" + t].edge_index
            srcs.append(ei[0] + n0[i])
            dsts.append(ei[1] + (g0 if b[t].scope == "local" else 0))
            if b[t].scope == "local":
                g0 += b[t].n_global
        G = g0 if blocks[0][t].scope == "local" else blocks[0][t].n_global
        dst = torch.cat(dsts)
        whole[t].id = torch.arange(G)
        whole[t].people = torch.bincount(dst, minlength=G)
        whole["agent", "attends_" + t, t].edge_index = torch.stack((torch.cat(srcs), dst))
    return whole, n0


def _block_worker(rank, world_size, port, out):
    from grad_june.partition import partition_from_blocks
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        kw = dict(seed=9, device="cpu", agents_per_super_area=2000, super_areas_per_region=5)
        mine = make_synthetic_world(N_BLOCK, block=(rank, world_size), **kw)
        local = partition_from_blocks(mine)
        part = local._gj_partition
        whole, n0 = _assemble([make_synthetic_world(N_BLOCK, block=(j, world_size), **kw) for j in range(world_size)])
        assert part.bounds == n0
        ref = partition_world(whole, rank, world_size, bounds=n0)
        rp = ref._gj_partition
        assert sum(part.n_boundary.values()) > 0
        for t in part.types:
            assert torch.equal(local["attends_" + t].edge_index, ref["attends_" + t].edge_index), t
            assert torch.equal(local[t]["people"], ref[t]["people"]), t
            assert part.n_boundary[t] == rp.n_boundary[t], t
            assert torch.equal(part.touch_lid[t], rp.touch_lid[t]) and torch.equal(part.touch_pos[t], rp.touch_pos[t]), t
            assert torch.equal(part.owned[t], rp.owned[t]), t
            assert torch.equal(part.attend[t], rp.attend[t]), t
        assert part.n_boundary["household"] == 0 and part.n_boundary["company"] > 0 and part.n_boundary["leisure"] > 0
        # commuting stays inside the home region: only companies of the regions cut by a block border are shared
        assert part.n_boundary["company"] < 0.5 * int(mine["company"].n_global)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_blockwise_world_matches_partition_of_the_whole(world_size):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_block_worker, args=(world_size, _free_port(), out), nprocs=world_size, join=True)
    assert sorted(out.keys()) == list(range(world_size))
