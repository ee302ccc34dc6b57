"""Quick timing probe (not the bench): fused step fwd+bwd on a synthetic world, CUDA-event timed."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gradabm-june_b200"))
import torch

from grad_june import GradJune, Timer, ops
from grad_june.default_config import default_parameters
from grad_june.runner import Runner
from grad_june.world import make_synthetic_world

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda:0"
params = default_parameters()
params["system"]["device"] = dev
params["policies"] = {}
params["timer"]["total_days"] = steps
t0 = time.time()
data = make_synthetic_world(n, seed=0, device=dev)
torch.cuda.synchronize()
print(f"world gen {time.time()-t0:.2f}s", flush=True)
data = Runner.get_data(params, data=data)
model = GradJune.from_parameters(params)
for net in model.infection_networks.networks.values():
    net.log_beta = torch.nn.Parameter(net.log_beta)
runner = Runner(model=model, data=data, timer=Timer.from_parameters(params), log_fraction_initial_cases=-2.0,
                save_path="/tmp/x", parameters=params)
t0 = time.time()
from grad_june.world import get_device_world
w = get_device_world(data, dev)
torch.cuda.synchronize()
print(f"csr build {time.time()-t0:.2f}s edges/agent {w.n_edges/n:.2f} groups/agent {w.n_groups/n:.3f} "
      f"small {w.small_groups.numel()} chunks {w.chunk_group.numel()} big {w.big_groups.numel()}", flush=True)
for it in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    with ops.philox_seed(1):
        results, _ = runner()
    e1.record()
    loss = results["cases_per_timestep"].sum() + results["deaths_per_timestep"].sum()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    f, b = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print(f"iter {it}: fwd {f:.1f} ms bwd {b:.1f} ms  -> {n*steps/((f+b)*1e-3)/1e9:.3f} G agent-steps/s "
          f"({(f+b)/steps:.2f} ms/step) cases {results['cases_per_timestep'][[0,-1]].tolist()} "
          f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)

# host-side issue cost (no sync inside): how long Python needs to enqueue one step
import time as _t
torch.cuda.synchronize()
runner.timer.reset(); runner.restore_initial_data(); runner.set_initial_cases()
torch.cuda.synchronize()
t0 = _t.perf_counter()
reds = []
with ops.philox_seed(1):
    for _ in range(steps):
        next(runner.timer)
        _, red = model.step(runner.data, runner.timer, age_bins=(0, 18, 65, 100))
        reds.append(red)
t1 = _t.perf_counter()
torch.cuda.synchronize()
t2 = _t.perf_counter()
loss = torch.stack(reds)[:, :2].sum()
torch.cuda.synchronize()
t3 = _t.perf_counter()
loss.backward()
t4 = _t.perf_counter()
torch.cuda.synchronize()
t5 = _t.perf_counter()
print(f"host issue per fwd step {1e3*(t1-t0)/steps:.3f} ms (gpu drained after {1e3*(t2-t1):.2f} ms); "
      f"host issue per bwd step {1e3*(t4-t3)/steps:.3f} ms (gpu drained after {1e3*(t5-t4):.2f} ms)")
