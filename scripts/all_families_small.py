"""Small fused run through all four kernel families (was meant for `compute-sanitizer --tool memcheck`, which is closed on this pool): 3 timesteps fwd+bwd on a 70 001-agent world (odd size:
unaligned tile starts and an array end that is not a multiple of 16 bytes) through the pipelined kernels, with
quarantine active, then the same through the register-batched and the reference-order kernels."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gradabm-june_b200"))
import torch
from grad_june import GradJune, Timer, _lib, ops
from grad_june.default_config import default_parameters
from grad_june.runner import Runner
from grad_june.world import make_synthetic_world

DEV = "cuda:0"
params = default_parameters()
params["system"]["device"] = DEV
params["timer"]["total_days"] = 3
params["infection_seed"]["log_fraction_initial_cases"] = -1.5
params["policies"] = {"quarantine": {"quarantine": {1: {"start_date": "2022-01-01", "end_date": "2023-01-01",
                                                         "stage_threshold": 4}}}}
torch.manual_seed(5)
data = Runner.get_data(params, data=make_synthetic_world(70_001, seed=6, device=DEV, agents_per_super_area=5000))
model = GradJune.from_parameters(params)
runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-1.5,
                save_path="/tmp/gj_memcheck", parameters=params)
for name, pipe, look, exact in (("pipelined", True, False, False), ("pipelined+lookahead", True, True, False),
                                ("register-batched", False, False, False), ("reference order", False, False, True)):
    _lib.pipeline_enable(pipe, lookahead=look)
    ops.EXACT_ORDER = exact
    for net in model.infection_networks.networks.values():
        net.log_beta = torch.nn.Parameter(torch.as_tensor(net.log_beta).detach().to(DEV))
    with ops.philox_seed(3):
        results, _ = runner()
    (results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum()).backward()
    torch.cuda.synchronize()
    print(name, results["cases_per_timestep"].tolist(), flush=True)
ops.EXACT_ORDER = False
print("done")
