#!/usr/bin/env python
"""Per-kernel share of a prefill + Euler-step launch list (the `ncu --metrics gpu__time_duration.sum --csv` pass of
tools/profile_step.py; run here, no GPU):
    python tools/launch_breakdown.py profiles/r02h_launches_prefill_plus_1step.csv > profiles/r02h_step_breakdown.txt
The list is cut where the denoising step starts (first kernel after the prefill's last GEMM / rope launch that belongs
to the small per-step MLPs)."""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        us = v / 1e3 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1e3
        rows.append((r[col["Kernel Name"]], us))
    # the step starts at the timestep sinusoid kernel (first launch of predict())
    cut = next((i for i, (k, _) in enumerate(rows) if "timestep_sinusoid" in k), 0)
    for title, part in (("prefill", rows[:cut]), ("one Euler step", rows[cut:])):
        if not part:
            continue
        tot = sum(u for _, u in part)
        print(f"{title}: {len(part)} launches, {tot:.1f} us under ncu (serialised, cold caches)")
        agg = OrderedDict()
        for k, u in part:
            k = k.split("(")[0][:60]
            n, t = agg.get(k, (0, 0.0))
            agg[k] = (n + 1, t + u)
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {k:62s} n={n:4d} total {t:9.1f} us  avg {t / n:7.1f} us  {100 * t / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
