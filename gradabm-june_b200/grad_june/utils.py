"""Host-side helpers: path alias, dates, age-binned tables, distribution factory, seeding.

Semantics follow the reference's ``grad_june/utils.py`` (17-94); the synthetic two-group world of
``utils.py:97-133`` lives in :mod:`grad_june.world` (``create_simple_connected_graph``) because it
needs the world container rather than torch_geometric.
"""
import datetime as _dt
import random
from copy import deepcopy
from pathlib import Path

import numpy as np
import torch
from torch import distributions as _dist

from .paths import grad_june_path


def read_path(path_str):
    """``@grad_june/<rest>`` resolves against the directory that holds the package."""
    path = Path(path_str)
    if path.parts and path.parts[0] == "@grad_june":
        return grad_june_path.joinpath(*path.parts[1:])
    return path


def read_date(date):
    if type(date) is str:
        return _dt.datetime.strptime(date, "%Y-%m-%d")
    if isinstance(date, _dt.date):
        return _dt.datetime.combine(date, _dt.datetime.min.time())
    raise TypeError("date must be a string or a datetime.date object")


def parse_age_probabilities(age_dict, fill_value=0):
    """Expand {"lo-hi": p} into a per-age list of 100 values (age a in [lo, hi) gets p).

    Same lookup rule as utils.py:47-72: ranges are ordered by their lower bound, laid out as a flat
    edge list [lo0, hi0, lo1, hi1, ...] with the value list [fill, p0, fill, p1, ..., fill], and age
    ``a`` reads the slot ``searchsorted(edges, a + 1)`` — so gaps give ``fill_value`` and overlapping
    ranges resolve exactly as the reference's binary search does.
    """
    lows, highs, probs = [], [], []
    for key, p in age_dict.items():
        lo, hi = key.split("-")
        lows.append(int(lo))
        highs.append(int(hi))
        probs.append(p)
    order = np.argsort(lows)
    edges, slots = [], [fill_value]
    for i in order:
        edges += [lows[i], highs[i]]
        slots += [np.array(probs)[i], fill_value]
    return [slots[int(np.searchsorted(edges, age + 1))] for age in range(100)]


def parse_distribution(spec, device):
    """{"dist": "LogNormal", "loc": .., "scale": ..} -> torch.distributions object (utils.py:75-83)."""
    kwargs = deepcopy(spec)
    cls = getattr(_dist, kwargs.pop("dist"))
    return cls(**{k: torch.tensor(v, device=device, dtype=torch.float) for k, v in kwargs.items()})


def fix_seed(seed=None):
    if seed is None:
        seed = np.random.randint(0, 1000)
    print(f"Fixing seed to {seed}")
    for f in (torch.manual_seed, np.random.seed, random.seed, torch.cuda.manual_seed, torch.cuda.manual_seed_all):
        f(seed)
