"""Simulation driver (reference: grad_june/runner.py).

Builds model, world and timer from the YAML parameters, keeps a backup of the initial state so that
``runner()`` can be re-run for every calibration iteration, seeds the initial cases and loops the
fused step.  The per-step result reductions (cases, differentiable deaths, cases by age bin —
runner.py:167-171,198-224) are produced inside the step kernel, so the loop issues no extra passes
over the agents.
"""
import copy
from pathlib import Path

import numpy as np
import pandas as pd
import torch
import yaml

from . import ops
from .model import GradJune
from .partition import all_reduce_sum
from .paths import ensure_default_config
from .timer import Timer
from .transmission import PROFILE_KEYS, TransmissionSampler
from .utils import read_path
from .world import load_world, original_order, renumber_world

_STATE_KEYS = ("susceptibility", "is_infected", "infection_time", "transmission")
_SYMPTOM_KEYS = ("current_stage", "next_stage", "time_to_next_stage")


class Runner(torch.nn.Module):
    def __init__(self, model, data, timer, log_fraction_initial_cases, save_path, parameters,
                 age_bins=(0, 18, 65, 100)):
        super().__init__()
        self.model = model
        self.data = data
        self.data_backup = self.backup_infection_data(data)
        self.timer = timer
        self.log_fraction_initial_cases = log_fraction_initial_cases
        self.device = model.device
        self.age_bins = torch.tensor(age_bins, device=self.device)
        self._age_bins_host = tuple(int(b) for b in age_bins)
        self.ethnicities = np.sort(np.unique(data["agent"].ethnicity))
        self.n_agents = data["agent"].id.shape[0]
        self.population_by_age = self.get_people_by_age()
        self.save_path = Path(save_path)
        self.input_parameters = parameters
        # batched ensemble: ``runner.batch = b`` makes ``forward()`` step b independent samples of the
        # epidemic at once — the networks' ``log_beta`` then hold b values each, the state tensors are [b, Np] and
        # every result series is [T+1, b].  All samples draw the same noise (common random numbers).
        self.batch = None
        self.restore_initial_data()

    @classmethod
    def from_file(cls, fpath=None):
        with open(fpath or ensure_default_config(), "r") as f:
            return cls.from_parameters(yaml.safe_load(f))

    @classmethod
    def from_parameters(cls, params):
        return cls(
            model=GradJune.from_parameters(params),
            data=cls.get_data(params),
            timer=Timer.from_parameters(params),
            log_fraction_initial_cases=params["infection_seed"]["log_fraction_initial_cases"],
            save_path=params["save_path"],
            parameters=params,
            age_bins=params.get("age_bins_to_save", (0, 18, 65, 100)),
        )

    @staticmethod
    def get_data(params, data=None):
        """World + freshly sampled per-agent profile parameters + initial state (runner.py:65-91).
        ``data`` may be passed directly instead of ``params['data_path']`` (synthetic worlds)."""
        device = params["system"]["device"]
        if data is None:
            data = load_world(read_path(params["data_path"]))
        data = data.to(device)
        # Worlds arrive numbered as the reference's loaders number them (by area and age): renumber the agents
        # household-contiguous inside their leisure cell so that the streaming layout tiers apply
        # (world.layout_order).  ``data["agent"].original_index`` keeps the loaded numbering; the noise stream is
        # keyed by it and ``Runner.forward`` returns ``is_infected`` in it.  ``system: {renumber_agents: false}``
        # keeps the loaded numbering (one part of a partitioned world is never renumbered on its own).
        if params["system"].get("renumber_agents", True) and "_gj_partition" not in data.__dict__ \
                and "_gj_block" not in data.__dict__:
            data = renumber_world(data)
        n_agents = len(data["agent"]["id"])
        values = TransmissionSampler.from_parameters(params)(n_agents)
        if "original_index" in data["agent"] and "_gj_partition" not in data.__dict__:
            values = values[:, data["agent"]["original_index"]]     # drawn per ORIGINAL agent: independent of the layout
        data["agent"].infection_parameters = {key: values[i, :].contiguous() for i, key in enumerate(PROFILE_KEYS)}
        data["agent"].transmission = torch.zeros(n_agents, device=device)
        data["agent"].susceptibility = torch.ones(n_agents, device=device)
        data["agent"].is_infected = torch.zeros(n_agents, device=device)
        data["agent"].infection_time = torch.zeros(n_agents, device=device)
        data["agent"].symptoms = {
            "current_stage": torch.ones(n_agents, dtype=torch.long, device=device),
            "next_stage": torch.ones(n_agents, dtype=torch.long, device=device),
            "time_to_next_stage": torch.zeros(n_agents, device=device),
        }
        return data

    def backup_infection_data(self, data):
        agent = data["agent"]
        backup = {key: agent[key].detach().clone() for key in _STATE_KEYS}
        backup["symptoms"] = {key: agent["symptoms"][key].detach().clone() for key in _SYMPTOM_KEYS}
        return backup

    def restore_initial_data(self):
        agent = self.data["agent"]
        for key in _STATE_KEYS:
            agent[key] = self.data_backup[key].detach().clone()
        for key in _SYMPTOM_KEYS:
            agent.symptoms[key] = self.data_backup["symptoms"][key].detach().clone()
        self.data["results"] = {"deaths_per_timestep": None}

    def _restore_for_window(self):
        """``restore_initial_data`` for ``forward()``: the fused step never writes its inputs in place, so the window
        can START from the backup tensors themselves instead of clones of them, with the stage arrays already in the
        fp32 the kernels read (the reference's initial stages are int64 ones; same values).  At 56 M agents this
        saves ~5 GB of copies and dtype conversions per window.  The backup is re-read (identity check) if the user
        replaced it."""
        key = tuple(id(self.data_backup[k]) for k in _STATE_KEYS) + \
            tuple((id(self.data_backup["symptoms"][k]), self.data_backup["symptoms"][k]._version) for k in _SYMPTOM_KEYS)
        key = key + (self.batch,)
        hit = self.__dict__.get("_window_start")
        if hit is None or hit[0] != key:
            state = {k: self.data_backup[k].detach() for k in _STATE_KEYS}
            sym = {k: self.data_backup["symptoms"][k].detach().to(torch.float32) for k in _SYMPTOM_KEYS}
            if self.batch:      # b copies of the initial state, rows padded to a multiple of four agents
                n = self.n_agents
                n_pad = (n + 3) // 4 * 4

                def rows(t):
                    out = torch.zeros(self.batch, n_pad, dtype=torch.float32, device=t.device)
                    out[:, :n] = t.to(torch.float32)
                    return out

                state = {k: rows(v) for k, v in state.items()}
                sym = {k: rows(v) for k, v in sym.items()}
            hit = self.__dict__["_window_start"] = (key, state, sym)
        agent = self.data["agent"]
        for k in _STATE_KEYS:
            agent[k] = hit[1][k]
        agent.symptoms = dict(hit[2])
        self.data["results"] = {"deaths_per_timestep": None}

    def _fraction_tensor(self):
        fraction = 10.0 ** self.log_fraction_initial_cases
        dev = self.data["agent"].susceptibility.device
        if not torch.is_tensor(fraction):
            # a plain number: keep its device copy (no host-to-device copy per run; CUDA-graph capturable)
            key = (float(fraction), str(dev))
            if getattr(self, "_fraction_cache", (None, None))[0] != key:
                self._fraction_cache = (key, torch.tensor([float(fraction)], dtype=torch.float32).to(dev))
            return self._fraction_cache[1]
        return fraction.reshape(1).to(device=dev, dtype=torch.float32)

    def set_initial_cases(self):
        """infect_fraction_of_people + first symptoms update (runner.py:138-149) as one fused call."""
        _, red = self.model.step(self.data, self.timer, age_bins=self._age_bins_host, mode=ops.MODE_SEED,
                                 seed_fraction=self._fraction_tensor())
        return red

    def forward(self):
        timer, model, data = self.timer, self.model, self.data
        timer.reset()
        if self.data["agent"].susceptibility.is_cuda:
            self._restore_for_window()
        elif self.batch:
            raise RuntimeError("batched ensembles run on a CUDA device only")
        else:
            self.restore_initial_data()
        reds = [self.set_initial_cases()]
        dates = [timer.date]
        while timer.date < timer.final_date:
            next(timer)
            ahead = None
            if timer.date < timer.final_date:      # the schedule is known: tell the step what the next one will be
                ahead = copy.copy(timer)
                next(ahead)
            data, red = model.step(data, timer, age_bins=self._age_bins_host, want_probs=False, next_timer=ahead)
            reds.append(red)
            dates.append(timer.date)
        table = torch.stack(reds)                      # [T+1, 2 + n_bins]   (batched: [T+1, b, 2 + n_bins])
        table = all_reduce_sum(table, data.__dict__.get("_gj_partition"))   # partitioned world: sum over ranks
        ex = data.__dict__.get("_gj_cache", {}).get("exchange")
        if ex is not None and not torch.cuda.is_current_stream_capturing():
            ex.check()      # a peer that never arrived at an exchange must not go unnoticed (once per window)
        cases_per_timestep = table[..., 0]
        data["results"]["deaths_per_timestep"] = table[..., 1]
        results = {
            "dates": dates,
            "cases_per_timestep": cases_per_timestep,
            "daily_cases_per_timestep": torch.diff(
                cases_per_timestep, dim=0, prepend=torch.zeros_like(cases_per_timestep[:1])),
            "deaths_per_timestep": data["results"]["deaths_per_timestep"],
        }
        for i, key in enumerate(self._age_bins_host[1:]):
            results[f"cases_by_age_{key:02d}"] = table[..., 2 + i]
        is_infected = data["agent"].is_infected
        if is_infected.dim() == 2:                      # batched: [b, Np] -> [b, n_agents] in the loaded order
            is_infected = original_order(data, is_infected[:, : self.n_agents].t()).t()
            return results, is_infected
        return results, original_order(data, is_infected)

    def save_results(self, results, is_infected):
        self.save_path.mkdir(exist_ok=True, parents=True)
        df = pd.DataFrame(index=results["dates"])
        df.index.name = "date"
        for key, value in results.items():
            if key != "dates":
                df[key] = value.detach().cpu().numpy()
        df.to_csv(self.save_path / "results.csv")
        pd.DataFrame({"is_infected": is_infected.detach().cpu().numpy()}).to_csv(
            self.save_path / "results_is_infected.csv")

    # --- reference helper reductions, kept for API parity (runner.py:198-242) ----------------------
    def store_differentiable_deaths(self, data):
        symptoms = data["agent"].symptoms
        dead = self.model.symptoms_updater.stages_ids[-1]
        deaths = ((symptoms["current_stage"] == dead) * symptoms["current_stage"] / dead).sum()
        prev = data["results"]["deaths_per_timestep"]
        data["results"]["deaths_per_timestep"] = deaths if prev is None else torch.hstack((prev, deaths))

    def _age_mask(self, i):
        age = self.data["agent"].age
        return (age < self.age_bins[i]) * (age > self.age_bins[i - 1])

    def get_cases_by_age(self, data):
        ret = torch.zeros(self.age_bins.shape[0] - 1, device=self.device)
        for i in range(1, self.age_bins.shape[0]):
            ret[i - 1] = (data["agent"].is_infected * self._age_mask(i)).sum()
        return ret

    def get_people_by_age(self):
        return {int(self.age_bins[i].item()): self._age_mask(i).sum() for i in range(1, self.age_bins.shape[0])}

    def get_cases_by_ethnicity(self, data):
        """runner.py:226-242, in one pass on the device: the per-agent ethnicity code is built once (the reference
        rebuilds a boolean mask from the numpy strings per ethnicity and call), the sums are one index_add —
        differentiable with respect to is_infected like the reference's masked sums."""
        agent = data["agent"]
        eth = agent.ethnicity
        hit = self.__dict__.get("_ethnicity_codes")
        if hit is None or hit[0] is not eth or hit[2] != agent.is_infected.device:
            labels = np.asarray(eth)
            n = agent.is_infected.shape[0]
            if labels.shape[0] != n:             # synthetic worlds carry one label for everybody
                labels = np.broadcast_to(labels[:1], (n,))
            codes = np.searchsorted(self.ethnicities, labels)
            hit = self.__dict__["_ethnicity_codes"] = (eth, torch.as_tensor(codes, dtype=torch.long).to(
                agent.is_infected.device), agent.is_infected.device)
        ret = torch.zeros(len(self.ethnicities), device=agent.is_infected.device, dtype=agent.is_infected.dtype)
        return ret.index_add(0, hit[1], agent.is_infected)
