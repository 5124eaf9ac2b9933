#!/usr/bin/env python
"""Turn the raw artefacts of a gpurun call (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/collect_profiles.py r01

Reads (round r01): gpurun_out/bench_r1.json, launches.csv (ncu launch list of the bench command),
prof_tc_final.ncu-rep (ncu --set full of quantize_tc_kernel), side_ncu_*.csv, side_kernels.json.
Reads (later rounds): gpurun_out/<round>_bench_n1.json, <round>_launches.csv, <round>_prof_tc.ncu-rep, <round>_prof_side.ncu-rep.
Writes: profiles/<round>_*.{json,csv,txt} and profiles/search_kernel_traffic.json (read by bench.py).
"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    if hdr is None:
        return []
    ci = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[start:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        agg.setdefault(r[ci["Kernel Name"]].split("(")[0], []).append(float(r[ci["Metric Value"]].replace(",", "")))
    return [(k, len(v), sum(v) / len(v)) for k, v in agg.items()]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    summary = {"round": rnd}
    b = os.path.join(OUT, "bench_r1.json" if rnd == "r01" else f"{rnd}_bench_n1.json")
    if os.path.exists(b):
        shutil.copy(b, os.path.join(PROF, f"{rnd}_bench_line.json"))
        summary["bench"] = json.loads(open(b).read().strip().splitlines()[-1])
    l = os.path.join(OUT, "launches.csv" if rnd == "r01" else f"{rnd}_launches.csv")
    if os.path.exists(l):
        shutil.copy(l, os.path.join(PROF, f"{rnd}_ncu_launches.csv"))
        tab = launch_table(l)
        # (the one-off codebook preparation -- image, per-dimension sort -- is not part of a step)
        setup = ("DeviceRadixSort", "column_keys", "codebook_prep")
        per_step = [(k, n, t) for k, n, t in tab if n >= 5 and not any(x in k for x in setup)]
        tot = sum(t for _, _, t in per_step)
        summary["launch_shares"] = [{"kernel": k, "launches": n, "avg_us": t / 1e3, "share_of_step": t / tot} for k, n, t in per_step]
    rep = os.path.join(OUT, "prof_tc_final.ncu-rep" if rnd == "r01" else f"{rnd}_prof_tc.ncu-rep")
    if os.path.exists(rep):
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(PROF, f"{rnd}_search_kernel_ncu_details.txt"), "w").write(det)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active",
                "sm__inst_executed_pipe_alu", "sm__mem_tensor_cycles_active", "launch__registers_per_thread", "sm__warps_active",
                "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg")
        picked = {h: (vals[i], units[i]) for i, h in enumerate(hdr) if any(h.startswith(k) or k in h for k in keep) and "per_second" not in h}
        json.dump(picked, open(os.path.join(PROF, f"{rnd}_search_kernel_ncu_raw.json"), "w"), indent=1)

        def num(name):
            for i, h in enumerate(hdr):
                if h == name:
                    v = float(vals[i].replace(",", ""))
                    u = units[i].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            return None
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            json.dump({"kernel": "quantize_tc_kernel<32,0>" if rnd == "r01" else "quantize_tc_kernel<32,0,1> (fp16 accumulators)", "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                       "source": f"ncu --set full, profiles/{rnd}_search_kernel_ncu_details.txt"},
                      open(os.path.join(PROF, "search_kernel_traffic.json"), "w"), indent=1)
    srep = os.path.join(OUT, f"{rnd}_prof_side.ncu-rep")
    if os.path.exists(srep):
        det = subprocess.run(["ncu", "-i", srep, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(PROF, f"{rnd}_side_kernels_ncu_details.txt"), "w").write(det)
    side = {}
    for fn in sorted(glob.glob(os.path.join(OUT, "side_ncu_*.csv"))) if rnd == "r01" else []:
        rows = list(csv.reader(open(fn)))
        hdr = None
        for i, r in enumerate(rows):
            if r and r[0] == "ID":
                hdr, start = r, i + 1
                break
        if hdr is None:
            continue
        ci = {h: i for i, h in enumerate(hdr)}
        d = defaultdict(lambda: defaultdict(list))
        for r in rows[start:]:
            if len(r) < len(hdr):
                continue
            d[r[ci["Kernel Name"]].split("(")[0]][r[ci["Metric Name"]]].append(float(r[ci["Metric Value"]].replace(",", "")))
        for k, v in d.items():
            side[k] = {m: sum(x) / len(x) for m, x in v.items()}
    if side:
        summary["side_kernels_ncu"] = side
    sk = os.path.join(OUT, "side_kernels.json")
    if os.path.exists(sk):
        summary["side_kernels_events"] = json.load(open(sk))
    json.dump(summary, open(os.path.join(PROF, f"{rnd}_summary.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
