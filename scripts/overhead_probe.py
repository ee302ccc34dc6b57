"""Per-step host overhead of the fused step: Runner() + backward() on a small world (GPU time negligible)."""
import sys, time, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gradabm-june_b200")); sys.path.insert(0, str(ROOT))
import torch
from grad_june import GradJune, Runner, Timer, ops
from grad_june.default_config import default_parameters
from grad_june.world import freeze_device_world, make_synthetic_world

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = 10
dev = "cuda:0"
p = default_parameters(); p["system"]["device"] = dev; p["policies"] = {}; p["timer"]["total_days"] = steps
p["infection_seed"]["log_fraction_initial_cases"] = -2.0; p["save_path"] = tempfile.gettempdir() + "/gj_probe"
torch.manual_seed(0)
data = Runner.get_data(p, data=make_synthetic_world(N, seed=0, device=dev))
freeze_device_world(data, dev)
model = GradJune.from_parameters(p)
keys = list(model.infection_networks.networks.keys())
runner = Runner(model=model, data=data, timer=Timer.from_parameters(p), log_fraction_initial_cases=-2.0,
                save_path=p["save_path"], parameters=p)
lb = torch.tensor([float(model.infection_networks.networks[k].log_beta) for k in keys], device=dev)

def window():
    leaves = []
    for i, k in enumerate(keys):
        leaf = lb[i].detach().clone().requires_grad_(True)
        model.infection_networks.networks[k].log_beta = leaf
        leaves.append(leaf)
    with ops.philox_seed(7):
        results, _ = runner()
    t1 = time.perf_counter()
    loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum()
    loss.backward()
    return t1

for _ in range(3):
    window()
torch.cuda.synchronize()
for _ in range(3):   # one window at a time from an empty queue: host time of forward / backward, then the drain
    s0 = time.perf_counter(); t1 = window(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"  window from empty queue: fwd host {(t1 - s0) / steps * 1e3:.3f}  bwd host {(t2 - t1) / steps * 1e3:.3f}  "
          f"drain {(t3 - t2) / steps * 1e3:.3f} ms/step")
import gc
gc.callbacks.append(lambda phase, info: print(f"    [gc {phase} gen{info['generation']} collected={info.get('collected')}] t={time.perf_counter():.4f}") if info['generation'] >= 1 else None)
for i in range(8):
    s0 = time.perf_counter(); t1 = window(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"  w{i}: fwd host {(t1 - s0) / steps * 1e3:.3f}  bwd host {(t2 - t1) / steps * 1e3:.3f} ms/step  t={s0:.4f}")
gc.callbacks.clear()
t0 = time.perf_counter(); tf = 0.0
for _ in range(5):
    s = time.perf_counter(); t1 = window(); tf += t1 - s
torch.cuda.synchronize()
t = time.perf_counter() - t0
print(f"N={N}: {t / 5 / steps * 1e3:.3f} ms per step (fwd+bwd), forward host part {tf / 5 / steps * 1e3:.3f} ms/step")
if len(sys.argv) > 2 and sys.argv[2] == "cprof":
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3):
        window()
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
elif len(sys.argv) > 2:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        window(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=35))
