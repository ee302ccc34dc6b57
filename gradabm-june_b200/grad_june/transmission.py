"""Infectiousness profile: per-agent parameter sampling (world load) and the per-step updater.

``TransmissionUpdater`` keeps the reference's module interface (grad_june/transmission.py:38-51) and
evaluates the gamma-pdf profile in one CUDA pass (``gj_transmission_forward``), with the
time-independent factor exp(-lgamma(shape)) cached per world.
"""
import torch
import yaml

from . import ops
from .paths import ensure_default_config
from .utils import parse_distribution

PROFILE_KEYS = ("max_infectiousness", "shape", "rate", "shift")


class TransmissionSampler:
    """Draws the four per-agent profile parameters once, at world load (transmission.py:8-35)."""

    def __init__(self, max_infectiousness, shape, rate, shift):
        self.max_infectiousness = max_infectiousness
        self.shape = shape
        self.rate = rate
        self.shift = shift

    def __call__(self, n):
        return torch.vstack([getattr(self, key).rsample((n,)) for key in PROFILE_KEYS])

    @classmethod
    def from_file(cls, fpath=None):
        with open(fpath or ensure_default_config(), "r") as f:
            return cls.from_parameters(yaml.safe_load(f))

    @classmethod
    def from_parameters(cls, params):
        device = params["system"]["device"]
        return cls(**{key: parse_distribution(spec, device=device) for key, spec in params["transmission"].items()})


def profile_tensors(data):
    """(maxinf, shape, rate, shift, k0) as fp32 device tensors; k0 cached until the dict is replaced."""
    ip = data["agent"]["infection_parameters"]
    cache = data.__dict__.setdefault("_gj_cache", {})
    key = tuple(id(ip[k]) for k in PROFILE_KEYS)
    hit = cache.get("profile")
    if hit is not None and hit[0] == key:
        return hit[1]
    vals = [ops._f32(ip[k]) for k in PROFILE_KEYS]
    out = (*vals, ops.profile_k0(vals[1]))
    cache["profile"] = (key, out)
    return out


def profile_packed(data):
    """The profile as one float4 per agent (throughput-mode kernels); cached like ``profile_tensors``."""
    vals = profile_tensors(data)
    cache = data.__dict__.setdefault("_gj_cache", {})
    hit = cache.get("prof4")
    if hit is not None and hit[0] is vals[4]:
        return hit[1]
    out = ops.profile_pack(*vals)
    cache["prof4"] = (vals[4], out)
    return out


class TransmissionUpdater(torch.nn.Module):
    def forward(self, data, timer):
        maxinf, shape, rate, shift, k0 = profile_tensors(data)
        return ops.transmission(timer.now, data["agent"].infection_time, data["agent"].is_infected,
                                maxinf, shape, rate, shift, k0)
