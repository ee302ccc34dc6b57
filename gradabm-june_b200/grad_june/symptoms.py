"""Symptom-stage state machine (reference: grad_june/symptoms.py).

The per-agent update (transition when the dwell time has elapsed, Bernoulli branch by (stage, age),
LogNormal / Normal dwell-time draw, differentiable stage carrying) runs in the SYMPTOMS phase of
``gj_step_forward`` without the reference's per-stage host synchronisations.
"""
import torch
import yaml

from . import ops
from .paths import ensure_default_config
from .utils import parse_age_probabilities, parse_distribution

_DIST_KINDS = {"LogNormal": 0, "Normal": 1}


class SymptomsSampler:
    def __init__(self, stages, stage_transition_probabilities, stage_transition_times, recovery_times, device):
        self.stages = stages
        self.stages_ids = torch.arange(0, len(stages))
        self.device = device
        self.stage_transition_probabilities = self._parse_stage_transition_probabilities(
            stage_transition_probabilities, device=device)
        self.stage_transition_times = self._parse_stage_times(stage_transition_times, device=device)
        self.recovery_times = self._parse_stage_times(recovery_times, device=device)
        self._host_times = (self._host_table(stage_transition_times), self._host_table(recovery_times))

    @classmethod
    def from_file(cls, fpath=None):
        with open(fpath or ensure_default_config(), "r") as f:
            return cls.from_parameters(yaml.safe_load(f))

    @classmethod
    def from_parameters(cls, params):
        return cls(**params["symptoms"], device=params["system"]["device"])

    def _parse_stage_transition_probabilities(self, stage_transition_probabilities, device):
        table = torch.zeros((len(self.stages), 100), device=device)
        for i, stage in enumerate(self.stages):
            if stage in stage_transition_probabilities:
                table[i] = torch.tensor(parse_age_probabilities(stage_transition_probabilities[stage]),
                                        dtype=torch.float32, device=device)
        return table

    def _parse_stage_times(self, stage_times, device):
        return {i: (parse_distribution(stage_times[stage], device) if stage in stage_times else None)
                for i, stage in enumerate(self.stages)}

    def _host_table(self, stage_times):
        """(kind, loc, scale) host floats per stage index, taken from the config (no device sync)."""
        out = {}
        for i, stage in enumerate(self.stages):
            spec = stage_times.get(stage)
            if spec is None:
                out[i] = None
                continue
            if spec["dist"] not in _DIST_KINDS:
                raise NotImplementedError(f"dwell-time distribution {spec['dist']} (supported: LogNormal, Normal)")
            out[i] = (_DIST_KINDS[spec["dist"]], float(spec["loc"]), float(spec["scale"]))
        return out

    @staticmethod
    def _dist_entry(dist):
        """(kind, loc, scale) of a dwell-time distribution OBJECT (utils.parse_distribution: LogNormal / Normal).
        Read from the object, not from the YAML it was built from, so that replacing or retuning
        ``stage_transition_times[i]`` / ``recovery_times[i]`` takes effect as it does in the reference."""
        if dist is None:
            return None
        name = type(dist).__name__
        if name not in _DIST_KINDS:
            raise NotImplementedError(f"dwell-time distribution {name} (supported: LogNormal, Normal)")
        return (_DIST_KINDS[name], float(dist.loc), float(dist.scale))

    def tables_key(self):
        """Changes whenever a distribution object or the probability table is replaced or edited in place."""
        dists = [d for table in (self.stage_transition_times, self.recovery_times) for d in table.values()]
        vers = tuple((id(d), getattr(d.loc, "_version", 0), getattr(d.scale, "_version", 0)) for d in dists if d is not None)
        p = self.stage_transition_probabilities
        return (id(p), p._version, vers)

    def tables(self, device):
        key = (self.tables_key(), str(device))
        hit = self.__dict__.get("_tables_cache")
        if hit is None or hit[0] != key:
            n = len(self.stages)
            trans = {i: self._dist_entry(self.stage_transition_times.get(i)) for i in range(n)}
            rec = {i: self._dist_entry(self.recovery_times.get(i)) for i in range(n)}
            self._host_times = (trans, rec)
            tabs = ops.SymptomsTables(n_stages=n, stage_prob=ops._f32(self.stage_transition_probabilities,
                                                                       torch.device(device)), trans=trans, rec=rec)
            # the cache holds the objects its key was made from (an id() can be reused once they are gone)
            hit = self.__dict__["_tables_cache"] = (key, tabs, (self.stage_transition_probabilities,
                                                               dict(self.stage_transition_times), dict(self.recovery_times)))
        return hit[1]

    # reference helpers kept for API parity (symptoms.py:65-80)
    def _get_need_to_transition(self, current_stage, time_to_next_stage, time):
        return (time >= time_to_next_stage) * (current_stage < len(self.stages) - 1)

    def _get_prob_next_symptoms_stage(self, ages, stages):
        return self.stage_transition_probabilities[stages, ages]

    def sample_next_stage(self, ages, current_stage, next_stage, time_to_next_stage, time):
        """symptoms.py:82-128 — one update without new infections."""
        from .infection import _agent_only_world
        from .world import DeviceWorld

        ops.require_cuda(time_to_next_stage, "time_to_next_stage")
        dev = time_to_next_stage.device
        n = ages.shape[0]
        base = _agent_only_world(n, dev)
        world = DeviceWorld(**{k: v for k, v in base.__dict__.items() if not k.startswith("_")})
        world.cls = ages.to(dev).to(torch.uint8).contiguous()
        spec = ops.StepSpec(now=float(time), dt=0.0, day_type=0, nets=[], quarantine=None,
                            phases=ops.PHASE_SYMPTOMS, want_reductions=False)
        state = {"cur": current_stage, "nxt": next_stage, "ttn": time_to_next_stage}
        out = ops.infection_step(ops.StepStatic(world=world, symptoms=self.tables(dev)), spec, None, state,
                                 n_in=torch.zeros(n, device=dev))
        return out["cur"], out["nxt"], out["ttn"]


class SymptomsUpdater(torch.nn.Module):
    def __init__(self, symptoms_sampler):
        super().__init__()
        if not isinstance(symptoms_sampler, SymptomsSampler):
            raise TypeError("symptoms_sampler must be an instance of SymptomsSampler.")
        self.symptoms_sampler = symptoms_sampler

    @classmethod
    def from_file(cls, fpath=None):
        fpath = fpath or ensure_default_config()
        try:
            with open(fpath, "r") as f:
                params = yaml.safe_load(f)
        except FileNotFoundError:
            raise FileNotFoundError(f"No file found at {fpath}.")
        except yaml.YAMLError:
            raise yaml.YAMLError(f"Invalid YAML file at {fpath}.")
        return cls.from_parameters(params)

    @classmethod
    def from_parameters(cls, params):
        return cls(symptoms_sampler=SymptomsSampler.from_parameters(params))

    def forward(self, data, timer, new_infected):
        try:
            symptoms = data["agent"].symptoms
        except (KeyError, AttributeError):
            raise KeyError("data must contain the 'agent' key.")
        if not all(key in symptoms for key in ["current_stage", "next_stage", "time_to_next_stage"]):
            raise KeyError(
                "symptoms must contain the 'current_stage', 'next_stage', and 'time_to_next_stage' keys.")
        from .world import get_device_world

        dev = symptoms["time_to_next_stage"].device
        ops.require_cuda(symptoms["time_to_next_stage"], "symptoms")
        world = get_device_world(data, dev)
        spec = ops.StepSpec(now=timer.now, dt=0.0, day_type=0, nets=[], quarantine=None,
                            phases=ops.PHASE_SYMPTOMS, want_reductions=False)
        state = {"cur": symptoms["current_stage"], "nxt": symptoms["next_stage"],
                 "ttn": symptoms["time_to_next_stage"]}
        out = ops.infection_step(ops.StepStatic(world=world, symptoms=self.symptoms_sampler.tables(dev)), spec, None,
                                 state, n_in=new_infected.to(torch.float32))
        symptoms["current_stage"] = out["cur"]
        symptoms["next_stage"] = out["nxt"]
        symptoms["time_to_next_stage"] = out["ttn"]
        return symptoms

    @property
    def stages_ids(self):
        return self.symptoms_sampler.stages_ids
