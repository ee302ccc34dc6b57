"""grad_june.june_world_loader (vectorised JUNE world -> edge lists) against a direct restatement of the reference's
per-person loops (june_world_loader/network_loader.py:14-44, leisure_loader.py:38-73, agent_data_loader.py:20-33) on
a random JUNE-layout file given as plain arrays (h5py is not part of the image)."""
from collections import defaultdict

import numpy as np
import torch

from grad_june import june_world_loader as JL


def _fake_june_file(n=3000, n_sa=12, seed=0):
    rng = np.random.default_rng(seed)
    sa = np.sort(rng.integers(0, n_sa, n))
    area = sa * 3 + rng.integers(0, 3, n)
    ids0 = rng.integers(0, n // 3, n)
    spec0 = np.where(rng.random(n) < 0.04, b"care_home", b"household")
    ids0 = np.where(spec0 == b"care_home", rng.integers(0, 5, n), ids0)
    kind = rng.choice([b"company", b"school", b"university", b"care_home", b"none"], n, p=[0.5, 0.2, 0.05, 0.02, 0.23])
    ids1 = np.select([kind == b"company", kind == b"school", kind == b"university", kind == b"care_home"],
                     [rng.integers(0, 200, n), rng.integers(0, 6, n), rng.integers(0, 2, n), rng.integers(0, 5, n)], 0)
    return {
        "population": {"id": np.arange(n), "age": rng.integers(0, 100, n), "sex": rng.choice([b"m", b"f"], n),
                       "ethnicity": rng.choice([b"A1", b"B2", b"C3"], n), "area": area, "super_area": sa,
                       "group_ids": np.stack([ids0, ids1], 1), "group_specs": np.stack([spec0, kind], 1)},
        "geography": {"area_name": np.array([f"E{i:05d}".encode() for i in range(3 * n_sa)]),
                      "area_socioeconomic_indices": rng.random(3 * n_sa),
                      "super_area_coordinates": np.stack([50 + 5 * rng.random(n_sa), -3 + 4 * rng.random(n_sa)], 1),
                      "super_area_id": np.arange(n_sa)},
        "households": {"id": np.arange(n // 3)}, "care_homes": {"id": np.arange(5)}, "companies": {"id": np.arange(200)},
        "schools": {"id": np.arange(6)}, "universities": {"id": np.arange(2)},
    }


def _reference_network(f, spec, columns):
    ret = defaultdict(list)                                  # network_loader.py:14-25
    for column in columns:
        gids = f["population"]["group_ids"][:, column]
        specs = f["population"]["group_specs"][:, column]
        for i, (gid, s) in enumerate(zip(gids, specs)):
            if s.decode() != spec:
                continue
            ret[gid].append(i)
    adj_i, adj_j = [], []                                    # network_loader.py:30-37
    for gid, people in ret.items():
        for person in people:
            adj_i.append(person)
            adj_j.append(gid)
    return np.array([adj_i, adj_j]), ret


def _reference_leisure(f, k):
    from sklearn.neighbors import BallTree
    coords = np.array([np.deg2rad(c) for c in f["geography"]["super_area_coordinates"]])
    ids = f["geography"]["super_area_id"]
    tree = BallTree(coords, metric="haversine")
    per_sa = {sid: list(np.where(f["population"]["super_area"] == sid)[0]) for sid in ids}
    src, dst, people = [], [], []
    for sid in ids:
        _, ind = tree.query(coords[sid].reshape(1, -1), k=k)
        members = []
        for sa in ind[0]:
            members += per_sa[sa]
        src += members
        dst += [sid] * len(members)
        people.append(len(members))
    return np.array([src, dst]), np.array(people)


def test_vectorised_loader_matches_the_reference_loops():
    f = _fake_june_file()
    data = JL.world_from_june_arrays(f, k_leisure=3)
    assert data.venue_types() == ["household", "care_home", "company", "school", "university", "leisure"]
    for spec, (plural, columns) in JL.NETWORK_SPECS.items():
        ref_edges, ref_groups = _reference_network(f, spec, columns)
        ei = data["attends_" + spec].edge_index.numpy()
        assert np.array_equal(ei, ref_edges), spec                       # same edges in the same ORDER
        people = [len(ref_groups[i]) for i in f[plural]["id"]]
        assert data[spec].people.tolist() == people, spec
        assert torch.equal(data["rev_attends_" + spec].edge_index, data["attends_" + spec].edge_index.flip(0))
    ref_edges, ref_people = _reference_leisure(f, 3)
    assert np.array_equal(data["attends_leisure"].edge_index.numpy(), ref_edges)
    assert np.array_equal(data["leisure"].people.numpy(), ref_people)
    agent = data["agent"]
    assert agent.sex.tolist() == [1 if s == b"f" else 0 for s in f["population"]["sex"]]
    assert agent.area[0] == f["geography"]["area_name"][f["population"]["area"][0]].decode()
    assert agent.socioeconomic_index.dtype == torch.int8 and int(agent.socioeconomic_index.min()) >= 1
    assert agent.ethnicity.dtype.kind == "U"


def test_loaded_world_renumbers_onto_the_streaming_tiers():
    """A world as the JUNE loader numbers it (by area: households scattered) goes through Runner.get_data's
    renumbering onto household = RANGE, leisure = CELL."""
    from grad_june import world as W
    f = _fake_june_file(n=6000, n_sa=8, seed=3)
    # households of at most 6 consecutive-ish people inside one super-area, as JUNE builds them
    sa = f["population"]["super_area"]
    rng = np.random.default_rng(1)
    hh = np.zeros(len(sa), dtype=np.int64)
    nxt = 0
    for s in range(8):
        members = rng.permutation(np.nonzero(sa == s)[0])
        sizes = rng.integers(1, 7, len(members))
        ends = np.cumsum(sizes)
        grp = np.searchsorted(ends, np.arange(len(members)), side="right")
        hh[members] = nxt + grp
        nxt += grp.max() + 1
    f["population"]["group_ids"][:, 0] = hh
    f["population"]["group_specs"][:, 0] = b"household"
    f["households"]["id"] = np.arange(nxt)
    data = JL.world_from_june_arrays(f, k_leisure=3)
    types = data.venue_types()
    n = 6000

    def tiers(d):
        dw = W.build_csr(n, types, {t: d["attends_" + t].edge_index for t in types}, {t: d[t]["people"] for t in types},
                         {t: len(d[t]["id"]) for t in types}, d["agent"].age, d["agent"].sex, 16, 1024, "cpu")
        return dict(zip(types, dw.type_tier))
    before = tiers(data)
    assert before["household"] == W.TIER_GENERIC
    data = W.renumber_world(data)
    after = tiers(data)
    assert after["household"] == W.TIER_RANGE and after["leisure"] == W.TIER_CELL
