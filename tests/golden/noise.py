"""Deterministic injected noise shared by the golden generator and the tests (numpy PCG64)."""
import numpy as np


def make_noise(seed: int, n_calls: int, n_agents: int, n_stages: int = 8):
    """Per sampler/symptoms call c: E[2,N] ~ Exp(1), u[N] ~ U[0,1), z[2*(S-3),N] ~ N(0,1); float32."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_calls):
        E = rng.standard_exponential(size=(2, n_agents), dtype=np.float32)
        E = np.maximum(E, np.float32(1e-30))
        u = rng.random(size=(n_agents,), dtype=np.float32)
        z = rng.standard_normal(size=(2 * (n_stages - 3), n_agents), dtype=np.float32)
        out.append((E, u, z))
    return out
