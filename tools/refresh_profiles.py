#!/usr/bin/env python3
"""Copies what `tools/gpu_checks.sh` left in gpurun_out/ (r02_*) into profiles/ under the names profiles/README.md lists, with
the two warm launch lists condensed to kernel name + ns.
    python tools/refresh_profiles.py"""
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

COPIES = {
    "r02_bench_default.json": "r02_bench_default.json", "r02_bench_kernels.jsonl": "r02_bench_kernels.jsonl",
    "r02_k1_probe.jsonl": "r02_k1_probe.jsonl", "r02_k4_probe.jsonl": "r02_k4_probe.jsonl", "r02_icp_probe.jsonl": "r02_icp_probe.jsonl",
    "r02_knn_probe.jsonl": "r02_knn_probe.jsonl", "r02_k1_ncu_raw_subset.csv": "r02_k1_deproject_tma_ncu_raw_subset.csv",
    "r02_k1_nv12_ncu_raw_subset.csv": "r02_k1_nv12_packed8_ncu_raw_subset.csv", "r02_k34_ncu_raw_subset.csv": "r02_k3_k4_ncu_raw_subset.csv",
    "r02_icp_ncu_raw_subset.csv": "r02_icp_ncu_raw_subset.csv", "r02_k1_source.csv.gz": "r02_k1_source_page.csv.gz",
    "r02_clocks.csv": "r02_clocks_during_bench.csv", "r02_launches_bench_k1.csv": "r02_launches_bench_k1.csv",
}


def condensed(src):
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.reader(lines))
    kn, mv = rows[0].index("Kernel Name"), rows[0].index("Metric Value")
    return [(re.sub(r"\(.*$", "", r[kn]), r[mv].replace(",", "")) for r in rows[1:] if len(r) > mv]


def rewrite(dst, head_lines, rows):
    old = open(dst).read().splitlines()
    with open(dst, "w") as f:
        f.write("\n".join(old[:head_lines]) + "\n")
        for name, ns in rows:
            f.write(f'"{name}",{ns}\n')


def main():
    for src, dst in COPIES.items():
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
    with open(os.path.join(G, "r02_bench_reference.json")) as f:  # the script runs under `set -x`: keep the JSON line only
        line = [ln for ln in f if ln.startswith("{")][-1]
    with open(os.path.join(P, "r02_bench_reference.json"), "w") as f:
        f.write(line)
    icp = condensed(os.path.join(G, "r02_launches_icp_warm.csv"))
    last = max(i for i, (n, _) in enumerate(icp) if "k_knn_init" in n)  # the timed registration: from its index build on
    rewrite(os.path.join(P, "r02_launches_icp_warm.csv"), 4, icp[last:])
    k4 = [(n, v) for n, v in condensed(os.path.join(G, "r02_launches_k4_warm.csv")) if "k_vox" in n or "k_transform" in n or "k_bounds_init" in n]
    dst = os.path.join(P, "r02_launches_k4_warm.csv")
    kept = len([ln for ln in open(dst).read().splitlines() if ln.startswith('"')])
    rewrite(dst, 3, k4[-kept:])
    d = json.loads(open(os.path.join(P, "r02_bench_default.json")).read().strip().splitlines()[-1])
    print("value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]), "probe",
          round(d["e2e"]["copy_probe"]["value"]), "cpu", round(d["cpu_baseline"]["value"], 1), d["clocks"])
    for r in d.get("rows", []):
        print(f'{r["ms"]:.4f} ms  frac {r["roofline"]["frac"]:.3f}  {r["row"][:80]}')
    print("icp searches (us):", [int(float(v) / 1e3) for n, v in icp[last:] if "nn_search" in n])


if __name__ == "__main__":
    main()
