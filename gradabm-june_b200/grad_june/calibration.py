"""Batched evaluation of parameter samples (BASELINE config 5: a calibration ensemble of
log-beta samples on one world, sharded across the GPUs of a box).

The reference evaluates one sample per Python-driven run (example_scripts/run_model.py:6-11).  Here every rank holds
a replica of the world and its own captured window (:class:`grad_june.graphed.GraphedRunner`); a batch ``[B, K]`` of
log-beta vectors is dealt out round-robin, each rank replays its samples back to back (one graph launch per sample)
and the losses and gradients are all-gathered — no collective on the data path.
"""
from typing import Callable, Optional, Sequence

import torch

from .graphed import GraphedRunner


class EnsembleEvaluator:
    def __init__(self, runner, loss_fn: Callable[[dict], torch.Tensor], networks: Optional[Sequence[str]] = None,
                 seed: int = 0, process_group=None):
        import torch.distributed as dist

        self.graphed = GraphedRunner(runner, loss_fn, networks=networks, seed=seed)
        self.names = self.graphed.names
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.world_size = dist.get_world_size(process_group) if self.distributed else 1

    @torch.no_grad()
    def __call__(self, log_betas: torch.Tensor):
        """``log_betas``: [B, K] (same on every rank).  Returns (losses [B], d loss / d log_beta [B, K]) on every rank."""
        import torch.distributed as dist

        dev = self.graphed.device
        log_betas = log_betas.to(device=dev, dtype=torch.float32)
        B, K = log_betas.shape
        per = -(-B // self.world_size)                       # samples per rank, the last ones padded
        mine = torch.zeros(per, K + 1, device=dev)
        for j in range(per):
            i = j * self.world_size + self.rank               # round-robin: sample i goes to rank i % world_size
            if i >= B:
                break
            loss, grads, _ = self.graphed(log_betas[i])
            mine[j, 0] = loss
            mine[j, 1:] = grads
        if self.world_size > 1:
            every = [torch.empty_like(mine) for _ in range(self.world_size)]
            dist.all_gather(every, mine, group=self.group)
            table = torch.stack(every, dim=1).reshape(per * self.world_size, K + 1)[:B]   # [j, rank] -> i
        else:
            table = mine[:B]
        return table[:, 0].clone(), table[:, 1:].clone()
