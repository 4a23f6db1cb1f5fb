#!/usr/bin/env python3
"""Markdown tables from the JSON lines of `bench.py --workload msm|ntt` runs (BASELINE.json configs[4] sweep).
    python tools/sweep_table.py lines.jsonl > profiles/r2_sweep_....txt"""
import json
import sys

rows = []
for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            try:
                rows.append(json.loads(line))
            except ValueError:
                pass


def cpu(d, key, fmt="{:.1f}"):
    c = d.get("cpu_baseline")
    return fmt.format(c[key]) if c and key in c else "-"


for wl, unit in (("msm", "Mterms/s"), ("ntt", "Mpoints/s")):
    sel = [d for d in rows if d["metric"].startswith(wl)]
    if not sel:
        continue
    for curve in sorted({d["config"]["curve"] for d in sel}):
        print(f"\n## BLS12-{curve} {'G1 MSM' if wl == 'msm' else 'Fr NTT (forward, in place)'}: bench.py --workload {wl} --curve {curve} (steps timed after >= 3 warm-up steps)\n")
        if wl == "msm":
            print("| log2 n | GPUs | ms | Mterms/s | windows | accumulate: madd/s (frac of IMAD.WIDE peak) | whole MSM: frac of the n*W floor | CPU ms (cores) | speed-up | bit-exact vs CPU |")
            print("|---|---|---|---|---|---|---|---|---|---|")
        else:
            print("| log2 n | GPUs (replicas) | ms | Mpoints/s per GPU | GB/s algorithmic (frac of HBM) | frac of the IMAD.WIDE floor | CPU ms (cores) | speed-up | bit-exact vs CPU |")
            print("|---|---|---|---|---|---|---|---|---|")
        for d in sorted((d for d in sel if d["config"]["curve"] == curve), key=lambda d: (d["config"]["log_n"], d["n_gpus"])):
            c = d.get("cpu_baseline")
            sp = f"{c['cpu_ms'] / d['ms_per_step']:.0f}x" if c else "-"
            cpu_ms = f"{c['cpu_ms']:.0f} ({c['cores']})" if c else "-"
            be = str(c["bit_exact"]) if c else "-"
            r = d["roofline"]
            if wl == "msm":
                a = r["alu"]
                print(f"| {d['config']['log_n']} | {d['n_gpus']} | {d['ms_per_step']:.2f} | {d['value'] / 1e6:.1f} | {d['config'].get('windows', '-')} | "
                      f"{a['madds_per_s'] / 1e9:.2f} G ({a.get('frac', 0):.2f}) | {a.get('frac_whole_msm', 0):.2f} | {cpu_ms} | {sp} | {be} |")
            else:
                print(f"| {d['config']['log_n']} | {d['n_gpus']} | {d['ms_per_step']:.3f} | {d['value'] / d['n_gpus'] / 1e6:.0f} | {r['achieved']:.0f} ({r['frac']:.3f}) | "
                      f"{r['alu']['frac']:.2f} | {cpu_ms} | {sp} | {be} |")
