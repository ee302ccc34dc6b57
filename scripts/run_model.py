#!/usr/bin/env python
"""The reference's example_scripts/run_model.py against this package: load the YAML, make the household log_beta
a Parameter, run the window, back-propagate the total number of cases, save the results (runner.py:185-196).

    python scripts/run_model.py [config.yaml] [--calibrate N] [--target results.csv]

--calibrate N: N iterations of grad_june.calibration.Calibrator fitting the log-betas to the cases_per_timestep
column of --target (or, without a target, to the run's own output at the initial log-betas shifted by +0.2 — a
self-consistency exercise)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gradabm-june_b200"))

import torch

from grad_june import Runner

ap = argparse.ArgumentParser()
ap.add_argument("config", nargs="?", default=None)
ap.add_argument("--calibrate", type=int, default=0)
ap.add_argument("--target", default=None)
ap.add_argument("--lr", type=float, default=0.05)
args = ap.parse_args()

runner = Runner.from_file(args.config)
nets = runner.model.infection_networks.networks
if args.calibrate == 0:
    nets["household"].log_beta = torch.nn.Parameter(nets["household"].log_beta)
    results, is_infected = runner()
    cases = results["cases_per_timestep"].sum()
    cases.backward()
    print("cases", float(cases), "d cases / d log_beta_household", float(nets["household"].log_beta.grad))
    runner.save_results(results, is_infected)
else:
    from grad_june.calibration import Calibrator
    if args.target:
        import pandas as pd
        target = torch.tensor(pd.read_csv(args.target)["cases_per_timestep"].to_numpy(), dtype=torch.float32)
    else:
        with torch.no_grad():
            for net in nets.values():
                net.log_beta = net.log_beta + 0.2
            target = runner()[0]["cases_per_timestep"].detach().cpu()
            for net in nets.values():
                net.log_beta = net.log_beta - 0.2
    target = target.to(runner.data["agent"].susceptibility.device)
    scale = float(target.abs().max().clamp(min=1.0))
    cal = Calibrator(runner, loss_fn=lambda r: (((r["cases_per_timestep"] - target) / scale) ** 2).mean(), lr=args.lr)
    history = cal.fit(args.calibrate, callback=lambda row: print(f"{row['iteration']:4d}  loss {row['loss']:.6g}"))
    print("saved to", cal.save(history))
