"""Turn the outputs of scripts/ncu_profile.sh into the small text summaries committed under profiles/.

    python scripts/ncu_summarise.py gpurun_out/<tag> profiles/<name>

writes <name>_launches.md (per-kernel share of the step from the launch list), <name>_kernels.csv (key
`--set full` metrics of every captured launch) and <name>_stalls_<kernel>.txt (stall samples by source line)."""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]


def short(name):
    m = re.search(r"(k_[a-z_0-9]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")).replace("(bool)", "") if m else name[:40]


# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(src + "_launches.csv")) if len(r) > 14 and r[0].isdigit()]
names = [short(r[4]) for r in rows]
ns = [float(r[14]) for r in rows]
# last complete BPTT window = from the last seeding launch (k_agent_forward) to the end
starts = [i for i, n in enumerate(names) if n.startswith("k_agent_forward")]
lo = starts[-1] if starts else 0
agg = OrderedDict()
for n, t in zip(names[lo:], ns[lo:]):
    a = agg.setdefault(n, [0.0, 0])
    a[0] += t
    a[1] += 1
tot = sum(v[0] for v in agg.values())
with open(dst + "_launches.md", "w") as f:
    f.write(f"ncu launch list (gpu__time_duration.sum, --clock-control none), last BPTT window of `{src}_launches.csv`:\n"
            f"{len(names) - lo} launches, {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares, not absolutes)\n\n"
            "| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"| {n} | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {100 * t / tot:.1f}% |\n")

# ---- full capture: key metrics per launch --------------------------------------------------------
rep = src + "_lean_full.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr = {k: i for i, k in enumerate(rr[0])}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
with open(dst + "_kernels.csv", "w") as f:
    f.write("kernel," + ",".join(f"{w} [{rr[1][hdr[w]]}]" for w in want) + ",dram_GBps\n")
    for r in rr[2:]:
        vals = [r[hdr[w]] for w in want]
        gb = float(r[hdr["dram__bytes_read.sum"]]) + float(r[hdr["dram__bytes_write.sum"]])
        unit = rr[1][hdr["dram__bytes_read.sum"]]
        scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(unit, 1.0)
        us = float(r[hdr["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(rr[1][hdr["gpu__time_duration.sum"]], 1.0)
        f.write(short(r[hdr["Kernel Name"]]) + "," + ",".join(vals) + f",{gb * scale / (us * 1e-6):.0f}\n")

# ---- DRAM traffic per launch for bench.py's roofline record (roofline.traffic / traffic_source) ----------
import datetime
import json

ids = {"k_lean_transmission": "transmission", "k_lean_transmission_c": "transmission", "k_pipe_forward": "agent_forward", "k_lean_forward": "agent_forward",
       "k_pipe_backward_gather": "backward_gather", "k_lean_backward_gather": "backward_gather",
       "k_pipe_backward": "agent_backward", "k_lean_backward": "agent_backward", "k_lean_group_sums": "group_sums",
       "k_lean_group_fix": "group_fix", "k_lean_scatter_finalize": "group_small<fwd>", "k_lean_seed": "seeding"}
per = {}
for r in rr[2:]:
    base = re.sub(r"<.*", "", short(r[hdr["Kernel Name"]]))
    unit = rr[1][hdr["dram__bytes_read.sum"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    per.setdefault(ids.get(base, base), []).append(
        (float(r[hdr["dram__bytes_read.sum"]]) + float(r[hdr["dram__bytes_write.sum"]])) * scale)
# a full MIDDLE step on a WEEKDAY: the backward kernel's largest launches are the middle steps (see
# scripts/ncu_profile.sh), the group sums' largest launches the weekdays (companies and schools active)
kern = {k: (max(v) if k in ("agent_backward", "group_sums") else sum(v) / len(v)) for k, v in per.items()}
kern["group_chunk<bwd>"] = kern.get("group_sums", 0.0)
kern["group_fix<bwd>"] = kern.get("group_fix", 0.0)
n_agents = int(sys.argv[3]) if len(sys.argv) > 3 else 56_000_000
# forward: transmission (+ scatter), scatter finalize, agent forward; backward: agent backward, group sums + fix, gather
step = sum(kern.get(k, 0.0) for k in ("transmission", "group_small<fwd>", "agent_forward", "agent_backward", "group_sums",
                                      "group_fix", "backward_gather"))
git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump({"agents": n_agents, "git": git, "when": datetime.datetime.now(datetime.timezone.utc).isoformat(timespec="seconds"),
           "source": rep, "kernels": kern, "launches": {k: len(v) for k, v in per.items()},
           "step_bytes_per_agent": step / n_agents,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full); agent_backward = its "
                   "largest launch = a middle step of the window; group kernels run twice per step"},
          open(dst + "_traffic.json", "w"), indent=1)

# ---- stall samples by source line for the agent kernels ---------------------------------------------
for k in ("k_pipe_forward", "k_pipe_backward", "k_pipe_backward_gather", "k_lean_forward", "k_lean_backward",
          "k_lean_backward_gather", "k_lean_transmission", "k_lean_transmission_c", "k_lean_group_sums"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name-base", "function", "--kernel-name", "regex:^" + k + "$", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(out)
    txt = subprocess.run([sys.executable, "scripts/ncu_by_line.py", tmp, "30"], capture_output=True, text=True).stdout
    open(dst + "_stalls_" + k + ".txt", "w").write(txt)
print("written", dst + "_*")
